"""P2P transfers BETWEEN boxes: GPU (or host) tensors over the multipath datagram transport.

Inside a box the P2P engine maps the peer's memory over NVLink (CUDA IPC) and copies with a TMA kernel; a
peer on another box is not load/store reachable, so the bytes ride ``uccl_b200.net``: the sender stages GPU
data through pinned chunks (D2H of chunk k+1 overlaps the network send of chunk k), the receiver posts one
receive per chunk into pinned memory and copies each chunk up as it completes.  This is the role of the
reference's RDMA data path for cross-node KV-cache transfer (p2p/rdma/*, p2p/engine.cc ``send/recv`` and the
NIXL backend); with a GPUDirect packet backend under ``net.Engine`` the staging copies would disappear while
this API stays.

    # decode node                                   # prefill node
    ch = NetChannel.listen(engine); addr = ch.address    ch = NetChannel.connect(engine, addr)
    ch.accept()                                          ch.send_tensors(kv_blocks)
    ch.recv_tensors(kv_blocks)
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from ..net import Engine


class NetChannel:
    """One bidirectional flow of a :class:`uccl_b200.net.Engine` with tensor-level send / recv."""

    def __init__(self, engine: Engine, flow: Optional[int] = None, listen_id: Optional[int] = None,
                 chunk_bytes: int = 4 << 20, timeout_ms: int = 120000, force_staging: bool = False):
        self.engine, self.flow, self._lid = engine, flow, listen_id
        self.chunk_bytes, self.timeout_ms = chunk_bytes, timeout_ms
        self.force_staging = force_staging  # run host tensors through the chunked staging pipeline too (tests)
        self._pin: List[torch.Tensor] = []

    # ---- connection
    @classmethod
    def listen(cls, engine: Engine, **kw) -> "NetChannel":
        return cls(engine, listen_id=engine.listen(), **kw)

    @property
    def address(self) -> Tuple[str, int, int]:
        """(ip, port, listen_id): hand it to the peer through any control channel."""
        return (self.engine.address, self.engine.port, self._lid)

    def accept(self) -> "NetChannel":
        self.flow = self.engine.accept(self._lid, self.timeout_ms)
        return self

    @classmethod
    def connect(cls, engine: Engine, address: Sequence, **kw) -> "NetChannel":
        ip, port, lid = address
        ch = cls(engine, **kw)
        ch.flow = engine.connect(ip, int(port), int(lid), ch.timeout_ms)
        return ch

    def close(self) -> None:
        if self.flow is not None:
            self.engine.close(self.flow)
            self.flow = None

    # ---- staging
    def _pinned(self, i: int) -> torch.Tensor:
        while len(self._pin) <= i:
            self._pin.append(torch.empty(self.chunk_bytes, dtype=torch.uint8, pin_memory=torch.cuda.is_available()))
        return self._pin[i]

    @staticmethod
    def _expect(got: int, want: int) -> None:
        if got != want:
            raise RuntimeError(f"uccl_b200.p2p: received a {got}-byte chunk where {want} bytes were expected "
                               "(tensor size or chunk_bytes differ between the two ends)")

    @staticmethod
    def _record(stream):
        if stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record(stream)
        return ev

    @staticmethod
    def _bytes(t: torch.Tensor) -> torch.Tensor:
        if not t.is_contiguous():
            raise ValueError("uccl_b200.p2p: tensors must be contiguous")
        return t.view(-1).view(torch.uint8)

    # ---- data
    def send_tensor(self, t: torch.Tensor) -> int:
        """Blocking send of one tensor (any device).  Returns the number of bytes."""
        b = self._bytes(t)
        n = b.numel()
        nchunks = max(1, (n + self.chunk_bytes - 1) // self.chunk_bytes)
        if not t.is_cuda and not self.force_staging:
            # host memory goes out in place; the wire format (one message per chunk) is device independent
            ws = [self.engine.isend(self.flow, b[k * self.chunk_bytes: min((k + 1) * self.chunk_bytes, n)])
                  for k in range(nchunks)]
            for w in ws:
                w.wait(self.timeout_ms)
            return n
        stream = torch.cuda.current_stream(t.device) if t.is_cuda else None
        works, events = [None, None], [None, None]

        def stage(k):
            lo = k * self.chunk_bytes
            hi = min(lo + self.chunk_bytes, n)
            h = self._pinned(k % 2)[: hi - lo]
            h.copy_(b[lo:hi], non_blocking=True)
            events[k % 2] = (self._record(stream), h)

        stage(0)
        for k in range(nchunks):
            ev, h = events[k % 2]
            if ev is not None:
                ev.synchronize()
            works[k % 2] = self.engine.isend(self.flow, h)
            if k + 1 < nchunks:
                if works[(k + 1) % 2] is not None:  # the other pinned buffer must have left the host
                    works[(k + 1) % 2].wait(self.timeout_ms)
                stage(k + 1)
        for w in works:
            if w is not None:
                w.wait(self.timeout_ms)
        return n

    def recv_tensor(self, t: torch.Tensor) -> int:
        """Blocking receive into ``t`` (same byte size as the matching send)."""
        b = self._bytes(t)
        n = b.numel()
        nchunks = max(1, (n + self.chunk_bytes - 1) // self.chunk_bytes)
        if not t.is_cuda and not self.force_staging:
            ws = [(self.engine.irecv(self.flow, b[k * self.chunk_bytes: min((k + 1) * self.chunk_bytes, n)]),
                   min((k + 1) * self.chunk_bytes, n) - k * self.chunk_bytes) for k in range(nchunks)]
            for w, want in ws:
                self._expect(w.wait(self.timeout_ms), want)
            return n
        depth = 4  # receives posted ahead: data lands in place without waiting for a round trip per chunk
        stream = torch.cuda.current_stream(t.device) if t.is_cuda else None
        posted, copied_ev = {}, {}

        def post(k):
            slot = 2 + k % depth
            ev = copied_ev.pop(slot, None)  # the H2D copy that last read this pinned buffer must be done
            if ev is not None:
                ev.synchronize()
            lo = k * self.chunk_bytes
            hi = min(lo + self.chunk_bytes, n)
            h = self._pinned(slot)[: hi - lo]
            posted[k] = (self.engine.irecv(self.flow, h), h, lo, hi, slot)

        for k in range(min(depth, nchunks)):
            post(k)
        for k in range(nchunks):
            w, h, lo, hi, slot = posted.pop(k)
            self._expect(w.wait(self.timeout_ms), hi - lo)
            b[lo:hi].copy_(h, non_blocking=True)
            copied_ev[slot] = self._record(stream)
            if k + depth < nchunks:
                post(k + depth)
        if stream is not None:
            stream.synchronize()
        return n

    def send_tensors(self, ts: Sequence[torch.Tensor]) -> int:
        """Vectorised send (e.g. the KV blocks of one request); a small manifest goes first so that the
        receiver can check it is getting what it expects."""
        manifest = torch.tensor([len(ts)] + [t.numel() * t.element_size() for t in ts], dtype=torch.int64)
        self.engine.send(self.flow, manifest, self.timeout_ms)
        verdict = torch.zeros(1, dtype=torch.uint8)
        self.engine.recv(self.flow, verdict, self.timeout_ms)  # one round trip per request, not per block
        if int(verdict) != 1:
            raise RuntimeError("uccl_b200.p2p: the receiver rejected the transfer (it expects a different block layout)")
        return sum(self.send_tensor(t) for t in ts)

    def recv_tensors(self, ts: Sequence[torch.Tensor]) -> int:
        manifest = torch.zeros(len(ts) + 1, dtype=torch.int64)
        got = self.engine.recv(self.flow, manifest, self.timeout_ms)
        want = [len(ts)] + [t.numel() * t.element_size() for t in ts]
        ok = got == manifest.numel() * 8 and manifest.tolist() == want
        self.engine.send(self.flow, torch.tensor([1 if ok else 0], dtype=torch.uint8), self.timeout_ms)
        if not ok:
            raise RuntimeError(f"uccl_b200.p2p: peer announced {manifest.tolist()[: got // 8]}, receiver expected {want}")
        return sum(self.recv_tensor(t) for t in ts)
