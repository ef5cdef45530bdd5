"""Helpers that travel with the P2P engine in the reference (p2p/utils.py): descriptor-limit setup, a retrying TCP
connect, length-prefixed pickled messages for out-of-band exchange, and the closed-interval index `uccl.collective`
uses to find the registration that covers a tensor.  Same names and call shapes; the index is this library's
:class:`uccl_b200.utils.regions.RegionIndex` (sorted array + prefix maximum) instead of an external interval-tree
package.
"""
from __future__ import annotations

import pickle
import resource
import socket
import struct
import sys
import time
from typing import Any, Iterator, List, Optional, Tuple

from ..utils.regions import RegionIndex

__all__ = ["set_files_limit", "create_socket_and_connect", "send_obj", "recv_obj", "ClosedIntervalTree"]


def set_files_limit(verbose: bool = True) -> Tuple[int, int]:
    """Raise the soft descriptor limit to the hard one (one TCP connection per peer plus IPC handles add up).
    Returns the (soft, hard) pair in effect afterwards; never raises."""
    try:
        soft, hard = resource.getrlimit(resource.RLIMIT_NOFILE)
        if soft < hard:
            resource.setrlimit(resource.RLIMIT_NOFILE, (hard, hard))
            soft = hard
        if verbose:
            print(f"uccl_b200.p2p: descriptor limit soft={soft} hard={hard}", file=sys.stderr)
        return soft, hard
    except Exception as e:  # noqa: BLE001 - best effort, like the reference
        if verbose:
            print(f"uccl_b200.p2p: could not raise the descriptor limit: {e}", file=sys.stderr)
        return (-1, -1)


def create_socket_and_connect(host, port, max_retries=None, initial_delay=0.5, backoff=2, max_delay=10, timeout=None):
    """TCP connect with exponential back-off (the peer's listener may not be up yet).  `max_retries=None` retries
    forever; otherwise OSError after that many failed retries."""
    attempt, delay = 0, float(initial_delay)
    while True:
        s = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        try:
            if timeout:
                s.settimeout(timeout)
            s.connect((host, int(port)))
            s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            return s
        except OSError as e:
            s.close()
            attempt += 1
            if max_retries is not None and attempt > max_retries:
                raise OSError(f"could not connect to {host}:{port} after {max_retries} retries: {e}") from e
            time.sleep(delay)
            delay = min(delay * backoff, max_delay)


_LEN = struct.Struct("!Q")  # 8-byte length, network order (the reference's framing)


def _recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray(n)
    view, got = memoryview(buf), 0
    while got < n:
        k = sock.recv_into(view[got:], n - got)
        if k == 0:
            raise ConnectionError("socket closed while receiving")
        got += k
    return bytes(buf)


def send_obj(sock: socket.socket, obj: Any, *, protocol: int = pickle.HIGHEST_PROTOCOL) -> None:
    payload = pickle.dumps(obj, protocol=protocol)
    sock.sendall(_LEN.pack(len(payload)) + payload)


def recv_obj(sock: socket.socket) -> Any:
    (n,) = _LEN.unpack(_recv_exact(sock, _LEN.size))
    if n == 0:
        return None
    return pickle.loads(_recv_exact(sock, n))


class ClosedIntervalTree:
    """Intervals [start, end] with both ends included, each carrying a value; several intervals may share bounds.
    Queries return ``(start, end, data)`` tuples ordered by (start, end)."""

    def __init__(self):
        self._idx: RegionIndex = RegionIndex()

    def add(self, start: int, end: int, data: Any) -> None:
        if end < start:
            raise ValueError(f"Invalid closed interval: end ({end}) < start ({start})")
        self._idx.add(start, end - start + 1, data)

    def remove(self, start: int, end: int, data: Any = None) -> int:
        return self._idx.remove_matching(start, end - start + 1, data, any_value=data is None)

    @staticmethod
    def _closed(rows) -> List[Tuple[int, int, Any]]:
        return [(s, s + n - 1, v) for s, n, v in rows]

    def query_containing(self, query_start: int, query_end: int) -> List[Tuple[int, int, Any]]:
        return self._closed(self._idx.containing(query_start, query_end - query_start + 1))

    def query_overlap(self, query_start: int, query_end: int) -> List[Tuple[int, int, Any]]:
        return self._closed(self._idx.overlapping(query_start, query_end - query_start + 1))

    def query_exact_match(self, query_start: int, query_end: int, data: Any = None) -> List[Tuple[int, int, Any]]:
        rows = self._idx.matching(query_start, query_end - query_start + 1)
        return self._closed([r for r in rows if data is None or r[2] == data])

    def clear(self) -> None:
        self._idx.clear()

    def __iter__(self) -> Iterator[Tuple[int, int, Any]]:
        return iter(self._closed(list(self._idx)))

    def __len__(self) -> int:
        return len(self._idx)

    def __str__(self) -> str:
        return "\n".join(f"[{s}, {e}] -> {d}" for s, e, d in self)
