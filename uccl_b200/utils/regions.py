"""Index of registered memory regions with containment lookup.

Role of the reference's interval tree (p2p/utils.py:114-206, used by `uccl.collective` to find the registration
that covers a tensor, including a view into a larger registered buffer).  Registered regions rarely overlap and
lookups outnumber updates, so this is a sorted array searched with `bisect` plus a running prefix maximum of the
region ends -- O(log n) lookup for disjoint regions, correct (linear in the number of overlapping candidates) when
regions nest or overlap.
"""
from __future__ import annotations

import bisect
from typing import Generic, List, Optional, Tuple, TypeVar

V = TypeVar("V")


class RegionIndex(Generic[V]):
    def __init__(self):
        self._starts: List[int] = []
        self._items: List[Tuple[int, int, V]] = []  # (start, end, value), sorted by (start, end)
        self._max_end: List[int] = []                # prefix maximum of `end`

    def __len__(self) -> int:
        return len(self._items)

    def _rebuild(self, lo: int) -> None:
        m = self._max_end[lo - 1] if lo > 0 else -1
        del self._max_end[lo:]
        for _, end, _ in self._items[lo:]:
            m = max(m, end)
            self._max_end.append(m)

    def add(self, start: int, size: int, value: V) -> None:
        if size <= 0:
            raise ValueError("RegionIndex.add: empty region")
        end = start + size
        i = bisect.bisect_right(self._starts, start)
        while i > 0 and self._items[i - 1][0] == start and self._items[i - 1][1] > end:
            i -= 1
        self._starts.insert(i, start)
        self._items.insert(i, (start, end, value))
        self._rebuild(i)

    def remove(self, start: int, size: Optional[int] = None) -> Optional[V]:
        """Removes the region that starts at `start` (with that size if given); returns its value or None."""
        i = bisect.bisect_left(self._starts, start)
        while i < len(self._items) and self._items[i][0] == start:
            if size is None or self._items[i][1] == start + size:
                _, _, v = self._items.pop(i)
                self._starts.pop(i)
                self._rebuild(i)
                return v
            i += 1
        return None

    def find(self, start: int, size: int = 1) -> Optional[Tuple[int, int, V]]:
        """The smallest registered region that contains [start, start + size): (region start, region size, value)."""
        end = start + max(size, 1)
        i = bisect.bisect_right(self._starts, start) - 1
        best = None
        while i >= 0 and self._max_end[i] >= end:  # nothing at or before i reaches `end` once the prefix max drops
            s, e, v = self._items[i]
            if e >= end and (best is None or e - s < best[1]):
                best = (s, e - s, v)
            i -= 1
        return best

    def containing(self, start: int, size: int = 1) -> List[Tuple[int, int, V]]:
        """Every region that contains [start, start + size), in (start, end) order."""
        end = start + max(size, 1)
        i = bisect.bisect_right(self._starts, start) - 1
        out = []
        while i >= 0 and self._max_end[i] >= end:
            s, e, v = self._items[i]
            if e >= end:
                out.append((s, e - s, v))
            i -= 1
        out.reverse()
        return out

    def overlapping(self, start: int, size: int = 1) -> List[Tuple[int, int, V]]:
        """Every region that shares at least one byte with [start, start + size)."""
        end = start + max(size, 1)
        hi = bisect.bisect_left(self._starts, end)  # regions starting at or after `end` cannot overlap
        out = []
        i = hi - 1
        while i >= 0 and self._max_end[i] > start:
            s, e, v = self._items[i]
            if e > start:
                out.append((s, e - s, v))
            i -= 1
        out.reverse()
        return out

    def matching(self, start: int, size: int) -> List[Tuple[int, int, V]]:
        """Every region with exactly these bounds (several registrations may share them)."""
        i = bisect.bisect_left(self._starts, start)
        out = []
        while i < len(self._items) and self._items[i][0] == start:
            if self._items[i][1] == start + size:
                out.append((start, size, self._items[i][2]))
            i += 1
        return out

    def remove_matching(self, start: int, size: int, value=None, any_value: bool = True) -> int:
        """Removes every region with exactly these bounds (and this value unless `any_value`); returns how many."""
        i = bisect.bisect_left(self._starts, start)
        first, n = i, 0
        while i < len(self._items) and self._items[i][0] == start:
            if self._items[i][1] == start + size and (any_value or self._items[i][2] == value):
                self._items.pop(i)
                self._starts.pop(i)
                n += 1
            else:
                i += 1
        if n:
            self._rebuild(first)
        return n

    def clear(self) -> None:
        self._starts.clear()
        self._items.clear()
        self._max_end.clear()

    def __iter__(self):
        return iter([(s, e - s, v) for s, e, v in self._items])

    def exact(self, start: int) -> Optional[V]:
        i = bisect.bisect_left(self._starts, start)
        return self._items[i][2] if i < len(self._items) and self._items[i][0] == start else None

    def values(self) -> List[V]:
        return [v for _, _, v in self._items]
