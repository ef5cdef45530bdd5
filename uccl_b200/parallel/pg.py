"""``torch.distributed`` integration: a process-group backend named ``"uccl_b200"`` whose
collectives run on the native kernels, so ``dist.all_reduce`` / DDP / FSDP-style code uses this
library by changing one string::

    import uccl_b200.parallel.pg            # registers the backend
    dist.init_process_group("uccl_b200", rank=r, world_size=n, store=store)

Role in the reference: the NCCL drop-in path that ``examples/ddp_train.py`` exercises through
``NCCL_NET_PLUGIN`` (examples/ddp_run.sh:19-25) and the ukernel torch extension's ProcessGroup
(experimental/ukernel/py/ukernel_ccl/__init__.py:171-290).  The unique id travels through the
c10d store; on a GPU-less box the communicator falls back to the host backend so the same code
path is covered by CPU CI (world_size >= 2 over shm).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from .comm import Communicator

BACKEND_NAME = "uccl_b200"

_OP_NAMES = {
    dist.ReduceOp.SUM: "sum",
    dist.ReduceOp.PRODUCT: "prod",
    dist.ReduceOp.MAX: "max",
    dist.ReduceOp.MIN: "min",
    dist.ReduceOp.AVG: "avg",
}


def _op_name(op) -> str:
    try:
        return _OP_NAMES[op]
    except (KeyError, TypeError):
        for k, v in _OP_NAMES.items():
            if op == k:
                return v
        raise ValueError(f"uccl_b200: unsupported reduce op {op}")


class _Work(dist._Work if hasattr(dist, "_Work") else dist.Work):
    """Stream-ordered completion: the kernels were enqueued on the caller's current stream, so
    `wait()` has nothing to block on (same contract as NCCL's work.wait() on the current stream)."""

    def __init__(self, result):
        super().__init__()
        self._result = result
        self._fut = torch.futures.Future()
        self._fut.set_result(result)

    def wait(self, timeout=None):
        return True

    def is_completed(self):
        return True

    def is_success(self):
        return True

    def get_future(self):
        return self._fut

    def result(self):
        return self._result


class _StreamWork(_Work):
    """Work of a collective enqueued on the backend's communication stream (the way ProcessGroupNCCL does it):
    the call returned immediately, `wait()` orders the CALLER's current stream after the collective, and the
    CUDA-aware future lets DDP chain its bucket callbacks -- so the all-reduce of one gradient bucket runs
    while the backward pass of the next is still computing."""

    def __init__(self, done_event, result, fut):
        dist.Work.__init__(self) if not hasattr(dist, "_Work") else dist._Work.__init__(self)
        self._ev = done_event
        self._result = result
        self._fut = fut

    def wait(self, timeout=None):
        torch.cuda.current_stream().wait_event(self._ev)
        return True

    def is_completed(self):
        return self._ev.query()

    def is_success(self):
        return True

    def synchronize(self):
        self.wait()


class _AsyncWork(_Work):
    """Work of a collective that runs on the multi-box helper thread: `wait()` blocks until it is done and orders the
    caller's stream after it; the future completes from the helper thread, which is what lets DDP overlap the
    gradient all-reduce of one bucket with the backward pass of the next."""

    def __init__(self, mn_work, result):
        import threading

        dist.Work.__init__(self) if not hasattr(dist, "_Work") else dist._Work.__init__(self)
        self._w = mn_work
        self._result = result
        self._fut = torch.futures.Future()

        def finish():
            try:
                mn_work.wait()
                if mn_work._ev_out is not None:
                    mn_work._ev_out.synchronize()
                self._fut.set_result(result)
            except BaseException as e:  # noqa: BLE001
                self._fut.set_exception(e)

        threading.Thread(target=finish, daemon=True).start()

    def wait(self, timeout=None):
        self._w.wait()
        return True

    def is_completed(self):
        return self._w.is_completed()

    def is_success(self):
        return self._w.is_completed() and self._w._err is None


_GROUPS: List["ProcessGroupUCCL"] = []


def default_communicator():
    """The Communicator of the most recently created ``uccl_b200`` process group (e.g. to allocate DDP's
    gradient buckets in its symmetric heap: ``with default_communicator().use_mem_pool(): ddp = DDP(model)``)."""
    return _GROUPS[-1].comm if _GROUPS else None


class ProcessGroupUCCL(dist.ProcessGroup):
    def __init__(self, store, rank: int, world_size: int, timeout=None, heap_bytes: Optional[int] = None):
        super().__init__(rank, world_size)
        _GROUPS.append(self)
        self._rank, self._world = rank, world_size
        key = "uccl_b200/uid/%d" % int(os.environ.get("UCCL_B200_PG_SEQ", "0"))
        if rank == 0:
            store.set(key, Communicator.create_unique_id())
        uid = bytes(store.get(key))
        host = not torch.cuda.is_available()
        heap = heap_bytes or int(os.environ.get("UCCL_B200_PG_HEAP_MB", "128" if host else "2048")) << 20
        stage = (8 << 20) if host else (128 << 20)
        # ranks per NVLink domain: torchrun's LOCAL_WORLD_SIZE, or UCCL_B200_LOCAL_SIZE to override
        local = int(os.environ.get("UCCL_B200_LOCAL_SIZE", os.environ.get("LOCAL_WORLD_SIZE", str(world_size))))
        if 0 < local < world_size and world_size % local == 0:
            # the group spans several boxes: NVLink kernels inside a box, datagram rails between boxes
            from .multinode import MultiNodeCommunicator

            seq = os.environ.get("UCCL_B200_PG_SEQ", "0")
            self.comm = MultiNodeCommunicator.from_store(store, rank, world_size, local, prefix=f"uccl_b200/mn{seq}",
                                                         heap_bytes=heap, stage_bytes=stage, host=host)
            if os.environ.get("UCCL_B200_PG_ASYNC", "1") != "0":
                from .multinode import AsyncMultiNode

                self._async = AsyncMultiNode(self.comm)  # all-reduce returns a live Work: DDP overlaps buckets
        else:
            self.comm = Communicator.init(uid, rank, world_size, heap_bytes=heap, stage_bytes=stage, host=host)
            # CUDA: collectives run on a high-priority communication stream and return a live Work
            # (UCCL_B200_PG_ASYNC=0 enqueues them on the caller's stream instead)
            if not host and os.environ.get("UCCL_B200_PG_ASYNC", "1") != "0":
                self._comm_stream = torch.cuda.Stream(device=self.comm.device, priority=-1)

    # ---- required plumbing
    def getBackendName(self):
        return BACKEND_NAME

    def size(self):
        return self._world

    def rank(self):
        return self._rank

    def _prep(self, t: torch.Tensor) -> torch.Tensor:
        if not t.is_contiguous():
            raise ValueError("uccl_b200: tensors must be contiguous")
        return t

    # ---- collectives (signatures of c10d::ProcessGroup)
    _async = None
    _comm_stream = None

    def _on_comm_stream(self, tensors, fn, result):
        """Run `fn()` (which enqueues kernels) on the communication stream, ordered after the caller's stream."""
        cs = self._comm_stream
        dev = self.comm.device
        cur = torch.cuda.current_stream(dev)
        cs.wait_stream(cur)
        fut = torch.futures.Future(devices=[dev])
        with torch.cuda.stream(cs):
            fn()
            for t in tensors:
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    t.record_stream(cs)
            done = torch.cuda.Event()
            done.record(cs)
            fut.set_result(result)  # CUDA-aware future: records the completion on the communication stream
        return _StreamWork(done, result, fut)

    def _drain_async(self):
        """Collectives that do not go through the helper thread / communication stream must not overtake the ones
        that did (every rank issues them in the same order; the kernels match by launch order)."""
        if self._async is not None:
            self._async.submit(lambda: None).wait()
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.comm.device).wait_stream(self._comm_stream)

    def allreduce(self, tensors: List[torch.Tensor], opts=None):
        op = _op_name(opts.reduceOp) if opts is not None else "sum"
        if self._async is not None:
            ts = [self._prep(t) for t in tensors]

            def run():
                for t in ts:
                    self.comm.all_reduce(t, op)
                return ts

            return _AsyncWork(self._async.submit(run), tensors)
        if self._comm_stream is not None and all(t.is_cuda for t in tensors):
            ts = [self._prep(t) for t in tensors]
            return self._on_comm_stream(ts, lambda: [self.comm.all_reduce(t, op) for t in ts], tensors)
        for t in tensors:
            self.comm.all_reduce(self._prep(t), op)
        return _Work(tensors)

    def allreduce_coalesced(self, tensors, opts=None):
        return self.allreduce(tensors, opts)

    def broadcast(self, tensors: List[torch.Tensor], opts=None):
        self._drain_async()
        root = opts.rootRank if opts is not None else 0
        for t in tensors:
            self.comm.broadcast(self._prep(t), root=root)
        return _Work(tensors)

    def allgather(self, output_tensors: List[List[torch.Tensor]], input_tensors: List[torch.Tensor], opts=None):
        self._drain_async()
        for outs, inp in zip(output_tensors, input_tensors):
            flat = torch.empty((self._world,) + tuple(inp.shape), dtype=inp.dtype, device=inp.device)
            self.comm.all_gather(flat, self._prep(inp))
            for i, o in enumerate(outs):
                o.copy_(flat[i])
        return _Work(output_tensors)

    def _allgather_base(self, output: torch.Tensor, input: torch.Tensor, opts=None):
        self._drain_async()
        self.comm.all_gather(self._prep(output), self._prep(input))
        return _Work(output)

    def allgather_into_tensor_coalesced(self, outputs, inputs, opts=None):
        self._drain_async()
        for o, i in zip(outputs, inputs):
            self.comm.all_gather(self._prep(o), self._prep(i))
        return _Work(outputs)

    def reduce_scatter(self, output_tensors: List[torch.Tensor], input_tensors: List[List[torch.Tensor]], opts=None):
        self._drain_async()
        op = _op_name(opts.reduceOp) if opts is not None else "sum"
        for out, ins in zip(output_tensors, input_tensors):
            flat = torch.stack([self._prep(t) for t in ins]).contiguous()
            self.comm.reduce_scatter(self._prep(out), flat, op)
        return _Work(output_tensors)

    def _reduce_scatter_base(self, output: torch.Tensor, input: torch.Tensor, opts=None):
        self._drain_async()
        op = _op_name(opts.reduceOp) if opts is not None else "sum"
        self.comm.reduce_scatter(self._prep(output), self._prep(input), op)
        return _Work(output)

    def reduce_scatter_tensor_coalesced(self, outputs, inputs, opts=None):
        self._drain_async()
        op = _op_name(opts.reduceOp) if opts is not None else "sum"
        for o, i in zip(outputs, inputs):
            self.comm.reduce_scatter(self._prep(o), self._prep(i), op)
        return _Work(outputs)

    def reduce(self, tensors: List[torch.Tensor], opts=None):
        self._drain_async()
        op = _op_name(opts.reduceOp) if opts is not None else "sum"
        root = opts.rootRank if opts is not None else 0
        for t in tensors:
            self.comm.reduce(self._prep(t), root=root, op=op)
        return _Work(tensors)

    def alltoall_base(self, output: torch.Tensor, input: torch.Tensor, output_split_sizes, input_split_sizes,
                      opts=None):
        self._drain_async()
        if not output_split_sizes and not input_split_sizes:
            self.comm.all_to_all(self._prep(output), self._prep(input))
        else:
            row = input[0].numel() if input.dim() > 1 else 1
            sc = [int(s) * row for s in (input_split_sizes or [input.size(0) // self._world] * self._world)]
            rc = [int(s) * row for s in (output_split_sizes or [output.size(0) // self._world] * self._world)]
            self.comm.all_to_all_v(self._prep(output), self._prep(input), sc, rc)
        return _Work(output)

    def alltoall(self, output_tensors, input_tensors, opts=None):
        self._drain_async()
        inp = torch.stack([self._prep(t) for t in input_tensors]).contiguous()
        out = torch.empty_like(inp)
        self.comm.all_to_all(out, inp)
        for i, o in enumerate(output_tensors):
            o.copy_(out[i])
        return _Work(output_tensors)

    def barrier(self, opts=None):
        self._drain_async()
        self.comm.barrier()
        if not self.comm.is_host:
            torch.cuda.current_stream().synchronize()
        return _Work(None)

    def send(self, tensors, dst_rank, tag=0):
        self._drain_async()
        # native staged send/recv kernel (CUDA) / mailbox protocol (host); tags are not used, operations
        # towards one peer match in posting order like NCCL's
        for t in tensors:
            self.comm.send(self._prep(t), dst_rank)
        return _Work(tensors)

    def recv(self, tensors, src_rank, tag=0):
        self._drain_async()
        for t in tensors:
            self.comm.recv(self._prep(t), src_rank)
        return _Work(tensors)


def _create(store, rank, world_size, timeout=None):
    return ProcessGroupUCCL(store, rank, world_size, timeout)


def register() -> None:
    if BACKEND_NAME.upper() in getattr(dist.Backend, "backend_list", []) or hasattr(dist.Backend, BACKEND_NAME.upper()):
        return
    try:
        dist.Backend.register_backend(BACKEND_NAME, _create, devices=["cpu", "cuda"])
    except TypeError:
        dist.Backend.register_backend(BACKEND_NAME, _create)


register()
