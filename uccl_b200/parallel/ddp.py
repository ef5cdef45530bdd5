"""Data-parallel glue: a DDP communication hook that reduces each gradient bucket with the
native allreduce (fused 1/N averaging in the kernel epilogue -- no separate div kernel).

Reference usage being replaced: ``DDP(model)`` over NCCL with the UCCL net plugin
(examples/ddp_train.py:78-98).
"""
from __future__ import annotations

import torch

from .comm import Communicator


def _comm_stream(comm: Communicator):
    cs = getattr(comm, "_ddp_stream", None)
    if cs is None and not comm.is_host:
        cs = comm._ddp_stream = torch.cuda.Stream(device=comm.device, priority=-1)
    return cs


def _run_async(comm: Communicator, buf: torch.Tensor, fn):
    """Runs `fn()` on the communicator's high-priority side stream (ordered after the current stream) and
    returns a CUDA-aware future: DDP chains its bucket callbacks on it, so the reduction of one bucket
    overlaps the backward pass that fills the next."""
    cs = _comm_stream(comm)
    if cs is None:  # host backend: synchronous
        fn()
        fut = torch.futures.Future()
        fut.set_result(buf)
        return fut
    cs.wait_stream(torch.cuda.current_stream(comm.device))
    fut = torch.futures.Future(devices=[comm.device])
    with torch.cuda.stream(cs):
        fn()
        buf.record_stream(cs)
        fut.set_result(buf)
    return fut


def allreduce_hook(comm: Communicator):
    """``model.register_comm_hook(None, allreduce_hook(comm))``"""

    def hook(state, bucket):
        buf = bucket.buffer()
        return _run_async(comm, buf, lambda: comm.all_reduce(buf, "avg"))

    return hook


def bf16_compress_hook(comm: Communicator):
    """fp32 gradient buckets are reduced in bf16 on the wire; the cast back to fp32 is fused in
    the allreduce epilogue (bf16 in -> fp32 accumulate -> fp32 out)."""

    def hook(state, bucket):
        buf = bucket.buffer()

        def run():
            if buf.dtype != torch.float32 or (buf.numel() * 2) % 16 != 0:
                comm.all_reduce(buf, "avg")
            else:
                low = buf.to(torch.bfloat16)
                comm.all_reduce(low, "avg", out=buf)

        return _run_async(comm, buf, run)

    return hook


def wrap_ddp(model: torch.nn.Module, comm: Communicator, process_group=None, compress: bool = False, **ddp_kwargs):
    from torch.nn.parallel import DistributedDataParallel as DDP

    ddp = DDP(model, process_group=process_group, **ddp_kwargs)
    ddp.register_comm_hook(None, bf16_compress_hook(comm) if compress else allreduce_hook(comm))
    return ddp
