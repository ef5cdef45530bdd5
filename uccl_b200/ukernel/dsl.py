"""Author your own collective and run it on the persistent-worker executor.

The reference lets users describe algorithms in the MSCCL++ DSL, ships them as JSON plans and interprets them in one
generic kernel (experimental/lite/collective/execution_kernel.hpp:898, plan lookup nccl.cu:1401-1408).  Here the same
role is played by a small global-view builder whose programs run on the ukernel executor (``UkCommunicator``): a
:class:`Program` is, per rank and per lane, a sequence of the four tile operations of ``csrc/ukernel/uk_plan.h``

    copy(rank, dst, src, nbytes)            local copy
    reduce(rank, dst, src, src2, nbytes)    dst = src (op) src2, element type / operator chosen at run time
    send(src_rank, dst_rank, dst, src, n)   src_rank copies its `src` into dst_rank's `dst` and signals; the matching
                                            wait is placed in dst_rank's program at the same point
    h = isend(...); wait(h)                 the two halves separately (post both directions of an exchange, then wait)
    signal(src_rank, dst_rank)              a send without payload

over three buffers per rank -- ``In(off)``, ``Out(off)``, ``Scratch(off)`` (byte offsets, 16-byte aligned).  Operations
of one (rank, lane) execute in the order they were written; different lanes of a rank run concurrently (slice the data
over lanes).  Because ``send`` emits both halves at once, the order in which a program is written is itself a valid
schedule: a program that validates cannot deadlock.

    prog = recursive_doubling_allreduce(nranks=8, nbytes=1 << 20, elem_size=2, nlanes=4)
    prog.validate()                                   # structural checks (C++: uk_check_bounds + uk_validate)
    prog.simulate(ins, outs, op="sum")                # all ranks on host memory (C++ greedy simulator)
    prog.save("rd_allreduce_8.json"); Program.load(...)
    work = prog.run(uk_comm, tensor, out, op="sum")   # this rank's part on the device worker (or the host backend)
"""
from __future__ import annotations

import json
from typing import Dict, List, Optional, Sequence, Tuple

import torch

Ref = Tuple[str, int]


def In(off: int = 0) -> Ref:
    return ("in", int(off))


def Out(off: int = 0) -> Ref:
    return ("out", int(off))


def Scratch(off: int = 0) -> Ref:
    return ("scratch", int(off))


class Program:
    FORMAT = "uccl_b200.ukernel.program/1"

    def __init__(self, name: str, nranks: int, nlanes: int = 1, in_bytes: int = 0, out_bytes: int = 0,
                 scratch_bytes: int = 0, elem_size: int = 1):
        if nranks < 1 or nlanes < 1:
            raise ValueError("Program: nranks and nlanes must be positive")
        self.name, self.nranks, self.nlanes = name, int(nranks), int(nlanes)
        self.in_bytes, self.out_bytes, self.scratch_bytes = int(in_bytes), int(out_bytes), int(scratch_bytes)
        self.elem_size = int(elem_size)
        self.ops: List[List[dict]] = [[] for _ in range(self.nranks)]

    # ------------------------------------------------------------------ authoring
    def _emit(self, rank: int, op: dict) -> None:
        if not 0 <= rank < self.nranks:
            raise ValueError(f"Program {self.name}: rank {rank} out of range")
        self.ops[rank].append(op)

    def copy(self, rank: int, dst: Ref, src: Ref, nbytes: int, lane: int = 0) -> None:
        self._emit(rank, dict(kind="copy", lane=int(lane), dst=dst, src=src, bytes=int(nbytes)))

    def reduce(self, rank: int, dst: Ref, src: Ref, src2: Ref, nbytes: int, lane: int = 0) -> None:
        self._emit(rank, dict(kind="reduce", lane=int(lane), dst=dst, src=src, src2=src2, bytes=int(nbytes)))

    def send(self, src_rank: int, dst_rank: int, dst: Ref, src: Ref, nbytes: int, lane: int = 0) -> None:
        if src_rank == dst_rank:
            raise ValueError(f"Program {self.name}: send to self (use copy)")
        self._emit(src_rank, dict(kind="send", lane=int(lane), peer=int(dst_rank), dst=dst, src=src, bytes=int(nbytes)))
        self._emit(dst_rank, dict(kind="recv", lane=int(lane), peer=int(src_rank), bytes=int(nbytes)))

    def isend(self, src_rank: int, dst_rank: int, dst: Ref, src: Ref, nbytes: int, lane: int = 0):
        """The sending half only; returns a handle for :meth:`wait`, which places the receive in dst_rank's program.
        Lets both partners of an exchange post their sends before either waits.  Waits of one (src, dst, lane)
        must be written in the order of their sends."""
        if src_rank == dst_rank:
            raise ValueError(f"Program {self.name}: send to self (use copy)")
        self._emit(src_rank, dict(kind="send", lane=int(lane), peer=int(dst_rank), dst=dst, src=src, bytes=int(nbytes)))
        return (int(src_rank), int(dst_rank), int(lane), int(nbytes))

    def wait(self, handle) -> None:
        src_rank, dst_rank, lane, nbytes = handle
        self._emit(dst_rank, dict(kind="recv", lane=lane, peer=src_rank, bytes=nbytes))

    def signal(self, src_rank: int, dst_rank: int, lane: int = 0) -> None:
        self.send(src_rank, dst_rank, Scratch(0), Scratch(0), 0, lane)

    # ------------------------------------------------------------------ checking
    def _scratch(self) -> List[int]:
        return [self.scratch_bytes] * self.nranks

    def validate(self) -> None:
        """Raises ValueError unless every reference stays inside its buffer, every send has its receive (same lane,
        same order, same size) and lanes / peers are in range."""
        from . import _uk

        err = _uk().validate_ops(self.nlanes, self._scratch(), self.ops, self.in_bytes, self.out_bytes, self.elem_size)
        if err:
            raise ValueError(f"Program {self.name}: {err}")

    def simulate(self, ins: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], op: str = "sum") -> None:
        """Runs all ranks over host tensors (one `In` and one `Out` per rank; `outs[r]` may be `ins[r]` for in-place
        programs) with the C++ greedy simulator: every (rank, lane) runs as far as it can, so a rank that is far ahead
        of its peers -- the situation that exposes unsafe buffer reuse -- is what gets exercised."""
        from . import _uk
        from ..parallel.comm import dtype_code, op_code

        if len(ins) != self.nranks or len(outs) != self.nranks:
            raise ValueError("simulate: one input and one output per rank")
        for t in list(ins) + list(outs):
            if t.is_cuda or not t.is_contiguous():
                raise ValueError("simulate: contiguous host tensors only")
        for i, o in zip(ins, outs):
            if i.numel() * i.element_size() < self.in_bytes or o.numel() * o.element_size() < self.out_bytes:
                raise ValueError("simulate: tensor smaller than the program's buffer size")
        err = _uk().simulate_ops(self.nlanes, self._scratch(), self.ops, self.in_bytes, self.out_bytes,
                                 dtype_code(ins[0].dtype), op_code(op), [t.data_ptr() for t in ins],
                                 [t.data_ptr() for t in outs])
        if err:
            raise RuntimeError(f"Program {self.name}: {err}")

    # ------------------------------------------------------------------ execution
    def run(self, uk_comm, inp: torch.Tensor, out: Optional[torch.Tensor] = None, op: str = "sum", stream=None,
            symmetric: bool = False):
        """Executes this rank's part on ``uk_comm`` (a ``UkCommunicator`` of ``nranks`` ranks; every rank calls).
        ``symmetric=True``: `inp` / `out` are ``comm.empty`` tensors at the same heap offset everywhere and are used in
        place by the peers; otherwise both are staged through the heap.  Returns a ``UkWork``."""
        from . import UkWork
        from ..parallel.comm import dtype_code, op_code

        out = inp if out is None else out
        if uk_comm.world_size != self.nranks:
            raise ValueError(f"Program {self.name} is written for {self.nranks} ranks, "
                             f"the communicator has {uk_comm.world_size}")
        uk_comm._check(inp), uk_comm._check(out)
        ib, ob = inp.numel() * inp.element_size(), out.numel() * out.element_size()
        if ib < self.in_bytes or ob < self.out_bytes:
            raise ValueError(f"Program {self.name}: needs {self.in_bytes} / {self.out_bytes} bytes, got {ib} / {ob}")
        if inp.element_size() != self.elem_size and any(o["kind"] == "reduce" for o in self.ops[uk_comm.rank]):
            raise ValueError(f"Program {self.name}: reductions were laid out for {self.elem_size}-byte elements")
        t = uk_comm._u.run_custom(self.nlanes, self.scratch_bytes, self.ops[uk_comm.rank], inp.data_ptr(), self.in_bytes,
                                  out.data_ptr(), self.out_bytes, dtype_code(inp.dtype), op_code(op),
                                  uk_comm._stream(stream), bool(symmetric))
        return UkWork(uk_comm._u, t, out)

    # ------------------------------------------------------------------ (de)serialisation
    def to_dict(self) -> Dict:
        return dict(format=self.FORMAT, name=self.name, nranks=self.nranks, nlanes=self.nlanes, in_bytes=self.in_bytes,
                    out_bytes=self.out_bytes, scratch_bytes=self.scratch_bytes, elem_size=self.elem_size,
                    ranks=[[dict(o) for o in r] for r in self.ops])

    @classmethod
    def from_dict(cls, d: Dict) -> "Program":
        if d.get("format") != cls.FORMAT:
            raise ValueError(f"not a {cls.FORMAT} document")
        p = cls(d["name"], d["nranks"], d["nlanes"], d["in_bytes"], d["out_bytes"], d["scratch_bytes"],
                d.get("elem_size", 1))
        if len(d["ranks"]) != p.nranks:
            raise ValueError("program document: one op list per rank expected")
        for r, ops in enumerate(d["ranks"]):
            for o in ops:
                o = dict(o)
                for k in ("dst", "src", "src2"):
                    if k in o:
                        o[k] = (str(o[k][0]), int(o[k][1]))
                p.ops[r].append(o)
        return p

    def to_json(self) -> str:
        return json.dumps(self.to_dict())

    @classmethod
    def from_json(cls, text: str) -> "Program":
        return cls.from_dict(json.loads(text))

    def save(self, path: str) -> None:
        with open(path, "w") as f:
            f.write(self.to_json())

    @classmethod
    def load(cls, path: str) -> "Program":
        with open(path) as f:
            return cls.from_json(f.read())

    def num_ops(self) -> int:
        return sum(len(r) for r in self.ops)


# ---------------------------------------------------------------------- slicing helper
def lane_slices(nbytes: int, nlanes: int, align: int) -> List[Tuple[int, int]]:
    """[(offset, bytes)] per lane: contiguous slices whose boundaries are multiples of `align` (>= 16)."""
    align = max(16, align)
    while align % 16:
        align *= 2
    per = -(-nbytes // nlanes)
    per = -(-per // align) * align
    out = []
    for lane in range(nlanes):
        lo = min(nbytes, lane * per)
        hi = min(nbytes, lo + per)
        out.append((lo, hi - lo))
    return out


# ---------------------------------------------------------------------- programs that are NOT among the built-in plans
def recursive_doubling_allreduce(nranks: int, nbytes: int, elem_size: int = 4, nlanes: int = 1) -> Program:
    """All-reduce in log2(N) exchange rounds (built-ins: ring, 2(N-1) steps, and full-mesh two-shot, 2 steps): in
    round k every rank swaps its whole partial result with rank ^ 2**k and adds what it received.  Latency-optimal for
    small messages on any topology; In and Out may be the same tensor.  Needs a power-of-two rank count."""
    if nranks & (nranks - 1):
        raise ValueError("recursive doubling needs a power-of-two number of ranks")
    rounds = nranks.bit_length() - 1
    p = Program(f"recursive_doubling_allreduce_{nranks}", nranks, nlanes, nbytes, nbytes, max(rounds, 1) * _pad(nbytes),
                elem_size)
    slices = lane_slices(nbytes, nlanes, elem_size)
    for r in range(nranks):
        for lane, (off, n) in enumerate(slices):
            if n:
                p.copy(r, Out(off), In(off), n, lane)
    for k in range(rounds):
        slot = k * _pad(nbytes)  # one scratch slot per round: a partner that is a round ahead cannot overwrite it
        for lane, (off, n) in enumerate(slices):
            if not n:
                continue
            hs = [p.isend(r, r ^ (1 << k), Scratch(slot + off), Out(off), n, lane) for r in range(nranks)]
            for h in hs:  # both partners have posted their halves before either waits
                p.wait(h)
            for r in range(nranks):
                p.reduce(r, Out(off), Out(off), Scratch(slot + off), n, lane)
    return p


def binomial_broadcast(nranks: int, nbytes: int, root: int = 0, nlanes: int = 1) -> Program:
    """Broadcast along a binomial tree: log2(N) rounds, in round k the 2**k ranks that already hold the data each
    forward it to one more (the built-in plan lets the root write to everybody).  In-place on Out; the root's Out
    is initialised from its In."""
    p = Program(f"binomial_broadcast_{nranks}_root{root}", nranks, nlanes, nbytes, nbytes, 0, 1)
    slices = lane_slices(nbytes, nlanes, 16)
    for lane, (off, n) in enumerate(slices):
        if n:
            p.copy(root, Out(off), In(off), n, lane)
    have = 1
    while have < nranks:
        for v in range(have):  # virtual rank v (root = 0) feeds v + have
            if v + have < nranks:
                s, d = (v + root) % nranks, (v + have + root) % nranks
                for lane, (off, n) in enumerate(slices):
                    if n:
                        p.send(s, d, Out(off), Out(off), n, lane)
        have *= 2
    return p


def _pad(n: int) -> int:
    return -(-n // 256) * 256
