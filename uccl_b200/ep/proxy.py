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
    def __init__(self, comm: Communicator, capacity: int = 4096, start: bool = True):
        self.comm = comm
        self._p = _native.C().EpProxy(comm._c, int(capacity))
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


FifoProxy = Proxy
