"""GPU -> CPU command queue and CPU proxy (``uccl.ep.Proxy`` / ``FifoProxy`` role).

A kernel pushes 32-byte commands into a host-pinned ring (``d2h_push`` in
``csrc/ep/d2h_queue.cuh``); the proxy thread executes them: copy-engine writes into any peer's
symmetric heap (no SM time), a 64-bit remote add ordered after those writes ("put with signal"),
or a notification handed to Python.  The EP dispatch/combine kernels of this library do not need
it -- they address peers directly -- it is the service channel for device-initiated transfers.
"""
from __future__ import annotations

from typing import List, Tuple

import torch

from .. import _native
from ..parallel.comm import Communicator


class Proxy:
    def __init__(self, comm: Communicator, capacity: int = 4096, start: bool = True, rail=None):
        """``rail``: a dedicated :class:`uccl_b200.net.NetCommunicator` of this rank's rail-mates (the same local rank
        of every box).  With it, destination ranks are GLOBAL (box-major) ranks: commands for another box travel
        over the datagram transport to the rail-mate proxy, which applies them over NVLink -- the reference's
        "CPU proxy posts RDMA" path (ep/src/proxy.cpp)."""
        self.comm = comm
        self._p = _native.C().EpProxy(comm._c, int(capacity))
        self._rail = rail
        if rail is not None:
            flows = [rail.flows.get(k, 0) for k in range(rail.world_size)]
            self._p.attach_link(rail.engine._native, flows, rail.rank, rail.world_size, comm.world_size)
        if start:
            self._p.start()

    def _stream(self, stream=None) -> int:
        return (stream or torch.cuda.current_stream(self.comm.device)).cuda_stream

    # ---- device-side issue helpers (one-thread kernels; real users call d2h_push from their own kernels)
    def device_write(self, dst_rank: int, src: torch.Tensor, dst_offset: int, signal_offset: int = -1,
                     signal_value: int = 1, stream=None):
        """Ask, *from the GPU*, for ``src`` (a tensor in this rank's symmetric heap) to be copied to
        ``dst_offset`` of ``dst_rank``'s heap; optionally add ``signal_value`` to the 64-bit counter at
        ``signal_offset`` of the same peer once the copy has landed."""
        C = _native.C()
        nbytes = src.numel() * src.element_size()
        so = self.comm._c.heap_offset(src.data_ptr())
        self._p.issue_from_device(C.D2H_WRITE, dst_rank, 0, so, int(dst_offset), nbytes, 0, self._stream(stream))
        if signal_offset >= 0:
            self._p.issue_from_device(C.D2H_ATOMIC, dst_rank, 0, 0, int(signal_offset), 0, int(signal_value),
                                      self._stream(stream))

    def device_notify(self, tag: int, value: int, stream=None):
        self._p.issue_from_device(_native.C().D2H_NOTIFY, 0, int(tag), 0, 0, 0, int(value), self._stream(stream))

    def poll_notifications(self) -> List[Tuple[int, int]]:
        return list(self._p.poll_notifications())

    def drain(self, timeout_s: float = 30.0):
        self._p.drain(timeout_s)

    def queue_handle(self):
        """(ring, head, tail, ack, capacity) device pointers: pass to your own kernels as D2HQueueDev."""
        return self._p.queue_handle()

    def bench_throughput(self, blocks: int = 8, threads: int = 128, per_thread: int = 64) -> float:
        """Commands per second the GPU can push through the queue (all threads issue concurrently)."""
        return self._p.bench_throughput(blocks, threads, per_thread, self._stream())

    def bench_latency(self, iters: int = 1000) -> float:
        """Mean GPU -> CPU -> GPU round trip of one command, in microseconds."""
        return self._p.bench_latency(iters, self._stream())

    def stats(self) -> dict:
        return self._p.stats()

    def stop(self):
        self._p.stop()


class ProxyLink:
    """The network half of the proxy on host memory: ``put`` / ``add`` / ``notify`` towards the rail-mate proxy of
    another box, applied there in issue order (put-with-signal).  ``heaps[l]`` is the flat uint8 "symmetric heap"
    of local rank ``l`` of THIS box that remote writes may target.  GPU proxies get the same object through
    ``Proxy(comm, rail=...)``; this class is the CPU-only stand-in and the unit-test surface."""

    def __init__(self, rail, heaps: List[torch.Tensor]):
        assert all(h.dtype == torch.uint8 and h.is_contiguous() and not h.is_cuda for h in heaps)
        self._rail, self._heaps = rail, heaps
        flows = [rail.flows.get(k, 0) for k in range(rail.world_size)]
        self._l = _native.C().EpProxyLink(rail.rank, rail.world_size, rail.engine._native, flows,
                                          [h.data_ptr() for h in heaps], min(h.numel() for h in heaps))

    def put(self, dst_box: int, dst_local: int, dst_offset: int, src: torch.Tensor) -> None:
        assert src.is_contiguous() and not src.is_cuda
        self._l.put(dst_box, dst_local, int(dst_offset), src.data_ptr(), src.numel() * src.element_size())

    def add(self, dst_box: int, dst_local: int, dst_offset: int, value: int = 1) -> None:
        self._l.add(dst_box, dst_local, int(dst_offset), int(value))

    def notify(self, dst_box: int, a: int, b: int) -> None:
        self._l.notify(dst_box, int(a), int(b))

    def flush(self, timeout_ms: int = 30000) -> None:
        self._l.flush(timeout_ms)

    def stats(self) -> dict:
        return self._l.stats()


FifoProxy = Proxy
