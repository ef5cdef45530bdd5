"""ukernel: launch-free collectives executed by a persistent worker kernel.

Three layers (all native, see ``csrc/ukernel``):

* :class:`Worker` -- per-lane CPU->device FIFOs drained by one persistent CTA each
  (copy / reduce / signal / wait tasks); on a CPU-only machine the same FIFOs are drained by
  host threads.
* the **planner** (:func:`plan`, :func:`validate`, :func:`simulate`) -- tile DAGs for ring and
  full-mesh AllReduce, AllToAll, AllGather and Barrier.
* :class:`UkCommunicator` / :class:`ProcessGroup` -- collectives over a
  :class:`uccl_b200.Communicator`'s symmetric heap, ordered against torch streams with stream
  memory operations instead of kernel launches.
* :mod:`uccl_b200.ukernel.dsl` -- write your own collective (``Program``: copy / reduce / send on
  In / Out / Scratch), validate and simulate it on the CPU, ship it as JSON, run it on the same executor
  (the role of the reference's MSCCL++-DSL plans and their interpreter kernel).

Reference parity: ``experimental/ukernel`` (persistent kernel, CCL planner/executor, torch
ProcessGroup with ``all_reduce / all_to_all_single / barrier``: ``ukernel_ccl/__init__.py:171-290``).
"""
from __future__ import annotations

import enum
from typing import List, Optional, Sequence

import torch

from .. import _native
from ..parallel.comm import Communicator, dtype_code, op_code

COLLS = {"allreduce": 0, "alltoall": 1, "allgather": 2, "barrier": 3, "reduce_scatter": 4, "broadcast": 5}
ALGOS = {"auto": 0, "ring": 1, "fullmesh": 2}


def _uk():
    return _native.C().uk


def plan(coll: str, nbytes: int, nranks: int, rank: int, nlanes: int = 1, tile_bytes: int = 1 << 20,
         elem_size: int = 1, algo: str = "auto", root: int = 0):
    """Plan of one rank: ``(text, [op dict, ...])``."""
    return _uk().plan(COLLS[coll], int(nbytes), nranks, rank, nlanes, int(tile_bytes), elem_size, ALGOS[algo], root)


def validate(coll: str, nbytes: int, nranks: int, nlanes: int = 1, tile_bytes: int = 1 << 20, elem_size: int = 1,
             algo: str = "auto", root: int = 0) -> str:
    """Cross-rank structural validation of the plans of all ranks; '' when consistent."""
    return _uk().validate(COLLS[coll], int(nbytes), nranks, nlanes, int(tile_bytes), elem_size, ALGOS[algo], root)


def simulate(coll: str, ins: Sequence[torch.Tensor], outs: Sequence[torch.Tensor], op="sum", nlanes: int = 1,
             tile_bytes: int = 1 << 20, algo: str = "auto", root: int = 0) -> str:
    """Execute the plans of all ranks over CPU tensors with the reference scheduler (greedy: a rank
    runs ahead as far as its dependencies allow).  Returns '' on success, else the failure."""
    n = len(ins)
    assert len(outs) == n and all(t.device.type == "cpu" and t.is_contiguous() for t in list(ins) + list(outs))
    dt = ins[0].dtype
    if coll == "allreduce":
        nbytes = ins[0].numel() * ins[0].element_size()
    elif coll in ("alltoall", "reduce_scatter"):
        nbytes = ins[0].numel() * ins[0].element_size() // n
    else:
        nbytes = ins[0].numel() * ins[0].element_size()
    return _uk().simulate(COLLS[coll], nbytes, n, nlanes, int(tile_bytes), dtype_code(dt), op_code(op), ALGOS[algo],
                          [t.data_ptr() for t in ins], [t.data_ptr() for t in outs], root)


class Worker:
    """Raw task interface to the persistent worker (``device=-1``: host threads)."""

    def __init__(self, device: int = 0, nlanes: int = 2, timeout_ms: int = 20000, idle_us: int = -1,
                 start: bool = True):
        self._w = _uk().Worker(int(device), int(nlanes), int(timeout_ms), int(idle_us))
        self.nlanes = nlanes
        if start:
            self._w.start()

    def copy(self, lane: int, dst: int, src: int, nbytes: int) -> int:
        return self._w.push(lane, _uk().OP_COPY, dst=dst, src=src, bytes=int(nbytes))

    def reduce(self, lane: int, dst: int, a: int, b: int, nbytes: int, dtype: torch.dtype, op="sum") -> int:
        return self._w.push(lane, _uk().OP_REDUCE, dst=dst, src=a, src2=b, bytes=int(nbytes), dtype=dtype_code(dtype),
                            redop=op_code(op))

    def signal(self, lane: int, addr: int, value: int = 1) -> int:
        return self._w.push(lane, _uk().OP_SIGNAL, sig_addr=addr, sig_val=int(value))

    def wait_value(self, lane: int, addr: int, value: int) -> int:
        return self._w.push(lane, _uk().OP_WAIT, sig_addr=addr, sig_val=int(value))

    def wait(self, lane: int, ticket: int, timeout_s: float = 30.0):
        self._w.wait(lane, ticket, timeout_s)

    def wait_all(self, timeout_s: float = 30.0):
        self._w.wait_all(timeout_s)

    def done(self, lane: int, ticket: int) -> bool:
        return self._w.done(lane, ticket)

    def stats(self) -> dict:
        return self._w.stats()

    def stop(self):
        self._w.stop()

    @property
    def error(self) -> int:
        return self._w.error

    @property
    def kernel_launches(self) -> int:
        """How often the worker kernel was (re)launched: it quits after ``idle_us`` without work so that
        device-wide synchronisation, cudaFree and lazy module loads never wait on an idle worker."""
        return self._w.kernel_launches


class UkWork:
    """Handle of one enqueued collective (host-side wait; on CUDA the issuing stream is already
    ordered after the collective, so ``wait()`` is only needed before touching results on the host
    from another stream)."""

    def __init__(self, uk, ticket: int, result=None):
        self._uk, self._ticket, self._result = uk, ticket, result

    def is_completed(self) -> bool:
        return self._uk.test(self._ticket)

    def wait(self, timeout_s: float = 60.0):
        self._uk.wait(self._ticket, timeout_s)
        return True

    def result(self):
        return self._result


class UkCommunicator:
    """Collectives executed by the persistent worker over ``comm``'s symmetric heap.  Construction
    is collective (every rank, same arguments).  Data is staged through the heap in ``staging_bytes``
    segments unless the caller declares the tensors symmetric (``symmetric=True``: ``comm.empty`` buffers
    at the same heap offset on every rank).  The communicator's own regions may sit at different
    offsets on different ranks (fragmented heaps): their offsets are all-gathered at construction."""

    def __init__(self, comm: Communicator, nlanes: int = 4, tile_bytes: int = 1 << 20,
                 staging_bytes: int = 32 << 20):
        self.comm = comm
        self.rank, self.world_size = comm.rank, comm.world_size
        self._u = _uk().Comm(comm._c, int(nlanes), int(tile_bytes), int(staging_bytes))

    def _stream(self, stream) -> int:
        if self.comm.is_host:
            return 0
        return (stream or torch.cuda.current_stream(self.comm.device)).cuda_stream

    def _check(self, t: torch.Tensor):
        if not t.is_contiguous():
            raise ValueError("uccl_b200.ukernel: tensors must be contiguous")
        if t.device != self.comm.device:
            raise ValueError(f"uccl_b200.ukernel: tensor on {t.device}, communicator on {self.comm.device}")

    def all_reduce(self, tensor: torch.Tensor, op="sum", out: Optional[torch.Tensor] = None, algo: str = "auto",
                   stream=None, symmetric: bool = False) -> UkWork:
        """``symmetric=True``: tensor/out come from ``comm.empty`` at the same heap offset on every rank
        (same allocation order everywhere) and are used in place by the peers; otherwise staged."""
        out = tensor if out is None else out
        self._check(tensor), self._check(out)
        code = op_code(op)
        avg = code == 4
        t = self._u.all_reduce(tensor.data_ptr(), out.data_ptr(), tensor.numel(), dtype_code(tensor.dtype),
                               0 if avg else code, ALGOS[algo], self._stream(stream), bool(symmetric))
        if avg:
            if self.comm.is_host:
                self._u.wait(t, 60.0)
            out.div_(self.world_size)  # CUDA: the current stream is already ordered after the collective
        return UkWork(self._u, t, out)

    def all_to_all_single(self, out: torch.Tensor, inp: torch.Tensor, stream=None, symmetric: bool = False) -> UkWork:
        self._check(inp), self._check(out)
        if inp.numel() % self.world_size or out.numel() != inp.numel():
            raise ValueError("uccl_b200.ukernel: all_to_all_single needs equal splits")
        t = self._u.all_to_all(inp.data_ptr(), out.data_ptr(), inp.numel() // self.world_size, dtype_code(inp.dtype),
                               self._stream(stream), bool(symmetric))
        return UkWork(self._u, t, out)

    def all_gather_into_tensor(self, out: torch.Tensor, inp: torch.Tensor, stream=None,
                               symmetric: bool = False) -> UkWork:
        self._check(inp), self._check(out)
        if out.numel() != inp.numel() * self.world_size:
            raise ValueError("uccl_b200.ukernel: all_gather output must hold world_size * input elements")
        t = self._u.all_gather(inp.data_ptr(), out.data_ptr(), inp.numel(), dtype_code(inp.dtype), self._stream(stream),
                               bool(symmetric))
        return UkWork(self._u, t, out)

    def reduce_scatter_tensor(self, out: torch.Tensor, inp: torch.Tensor, op="sum", stream=None) -> UkWork:
        self._check(inp), self._check(out)
        if inp.numel() != out.numel() * self.world_size:
            raise ValueError("uccl_b200.ukernel: reduce_scatter input must hold world_size * out.numel() elements")
        code = op_code(op)
        avg = code == 4
        t = self._u.reduce_scatter(inp.data_ptr(), out.data_ptr(), out.numel(), dtype_code(inp.dtype), 0 if avg else code,
                                   self._stream(stream))
        if avg:
            if self.comm.is_host:
                self._u.wait(t, 60.0)
            out.div_(self.world_size)
        return UkWork(self._u, t, out)

    def broadcast(self, tensor: torch.Tensor, root: int = 0, stream=None) -> UkWork:
        self._check(tensor)
        t = self._u.broadcast(tensor.data_ptr(), tensor.data_ptr(), tensor.numel(), dtype_code(tensor.dtype), int(root),
                              self._stream(stream))
        return UkWork(self._u, t, tensor)

    def barrier(self, stream=None) -> UkWork:
        return UkWork(self._u, self._u.barrier(self._stream(stream)))

    def stats(self) -> dict:
        return self._u.stats()

    def stop(self):
        self._u.stop()


class ReduceOp(enum.IntEnum):
    """Operator names / values of the reference's ``ukernel_ccl.ReduceOp`` (ukernel_ccl/__init__.py:19-25)."""

    SUM = 1
    PRODUCT = 2
    MAX = 3
    MIN = 4
    BAND = 5


_REDUCE_NAMES = {ReduceOp.SUM: "sum", ReduceOp.PRODUCT: "prod", ReduceOp.MAX: "max", ReduceOp.MIN: "min"}


def _op_name(op) -> str:
    if isinstance(op, str):
        return {"product": "prod"}.get(op.lower(), op.lower())
    try:
        return _REDUCE_NAMES[ReduceOp(int(op))]
    except (KeyError, ValueError):
        raise ValueError(f"uccl_b200.ukernel: reduce op {op!r} is not supported (sum / prod / max / min)") from None


class Work:
    """Completion handle with the reference's two methods (``wait``, ``is_completed``)."""

    def __init__(self, inner):
        self._inner = inner

    def wait(self):
        return self._inner.wait()

    def is_completed(self) -> bool:
        return self._inner.is_completed()


class _DoneWork:
    def wait(self):
        return True

    def is_completed(self) -> bool:
        return True


class ProcessGroup:
    """The reference's ``ukernel_ccl.ProcessGroup`` surface (ukernel_ccl/__init__.py:171-290): ``all_reduce``,
    ``all_to_all_single`` (equal or explicit split sizes), ``barrier``, ``rank`` / ``world_size`` / ``gpu_id`` /
    ``backend``, ``same_host`` / ``peer_transport`` -- synchronous by default, ``async_op=True`` returns a ``Work`` --
    plus ``all_gather_into_tensor`` / ``reduce_scatter_tensor`` / ``broadcast``.  Built over an existing
    :class:`uccl_b200.Communicator` (the reference builds its own transport from rank / world size / exchanger
    address: :func:`init_process_group` does that from ``torch.distributed`` or the environment)."""

    def __init__(self, comm: Communicator, **kw):
        self._uk = UkCommunicator(comm, **kw)
        self._comm = comm

    @property
    def rank(self) -> int:
        return self._comm.rank

    @property
    def world_size(self) -> int:
        return self._comm.world_size

    @property
    def gpu_id(self) -> int:
        d = self._comm.device
        return -1 if self._comm.is_host else int(d.index or 0)

    @property
    def backend(self) -> str:
        return "ukernel"

    def same_host(self, peer_rank: int) -> bool:
        if not 0 <= int(peer_rank) < self.world_size:
            raise ValueError(f"peer rank {peer_rank} out of range")
        return True  # one communicator = one NVLink domain (boxes are joined by UkNetCommunicator)

    def peer_transport(self, peer_rank: int) -> str:
        self.same_host(peer_rank)
        if int(peer_rank) == self.rank:
            return "self"
        return "host-shm" if self._comm.is_host else "nvlink"

    def _finish(self, w, async_op: bool):
        if async_op:
            return Work(w)
        w.wait()
        return None

    def all_reduce(self, tensor, op=ReduceOp.SUM, async_op: bool = False, tile_bytes: Optional[int] = None,
                   num_flows: Optional[int] = None):
        """`tile_bytes` / `num_flows` are per-call knobs of the reference; here the tile size and the lane count are
        fixed when the group is built (``ProcessGroup(comm, tile_bytes=..., nlanes=...)``) and the arguments are
        accepted for source compatibility."""
        return self._finish(self._uk.all_reduce(tensor, _op_name(op)), async_op)

    def all_to_all_single(self, output, input, output_split_sizes=None, input_split_sizes=None, async_op: bool = False,
                          tile_bytes: Optional[int] = None, num_flows: Optional[int] = None):
        n = self.world_size
        isp = list(input_split_sizes) if input_split_sizes is not None else None
        osp = list(output_split_sizes) if output_split_sizes is not None else None
        even = lambda sp, t: sp is None or (len(sp) == n and len(set(sp)) == 1 and sp[0] * n == t.size(0))  # noqa: E731
        if even(isp, input) and even(osp, output):
            return self._finish(self._uk.all_to_all_single(output, input), async_op)
        # explicit (uneven) splits along dim 0: the native all-to-all-v kernel moves them (the worker's plans are
        # equal-split); ordered after the worker's queue by a barrier on the same stream
        if isp is None or osp is None or len(isp) != n or len(osp) != n:
            raise ValueError("all_to_all_single: give both split lists with one entry per rank")
        if sum(isp) != input.size(0) or sum(osp) != output.size(0):
            raise ValueError("all_to_all_single: split sizes do not add up to the tensors' first dimension")
        row = input[0].numel() if input.dim() > 1 and input.size(0) else 1
        self._uk.barrier().wait()
        self._comm.all_to_all_v(output.view(-1), input.reshape(-1), [s * row for s in isp], [s * row for s in osp])
        return Work(_DoneWork()) if async_op else None

    def all_gather_into_tensor(self, output, input, async_op: bool = False):
        return self._finish(self._uk.all_gather_into_tensor(output, input), async_op)

    def reduce_scatter_tensor(self, output, input, op=ReduceOp.SUM, async_op: bool = False):
        return self._finish(self._uk.reduce_scatter_tensor(output, input, _op_name(op)), async_op)

    def broadcast(self, tensor, src: int = 0, async_op: bool = False):
        return self._finish(self._uk.broadcast(tensor, src), async_op)

    def barrier(self, async_op: bool = False):
        return self._finish(self._uk.barrier(), async_op)

    def shutdown(self):
        self._uk.stop()


# ---- functional API on a default group (reference: ukernel_ccl/__init__.py:346-430)
_DEFAULT_GROUP: Optional[ProcessGroup] = None


def _comm_from_exchanger(rank: int, world_size: int, gpu_id: int, exchanger_ip: str, exchanger_port: int,
                         heap_bytes: int, stage_bytes: int):
    """A Communicator for a world that has no torch.distributed: rank 0 serves a key/value exchanger
    (`ukernel.p2p.Exchanger`, the reference's socket OOB) and publishes the 128-byte unique id through it."""
    from .p2p import Exchanger, _ExchangerClient

    server = Exchanger(exchanger_ip, exchanger_port) if rank == 0 else None
    x = _ExchangerClient(exchanger_ip, exchanger_port)
    try:
        if rank == 0:
            x.put("ukernel_ccl/uid", Communicator.create_unique_id())
        ok, uid = x.get("ukernel_ccl/uid", 120000)
        if not ok:
            raise RuntimeError("ukernel: rank 0 never published the communicator id")
        host = not torch.cuda.is_available() or gpu_id is None or int(gpu_id) < 0
        comm = Communicator.init(uid, rank, world_size, device=None if host else int(gpu_id), host=host,
                                 heap_bytes=heap_bytes, stage_bytes=stage_bytes)
        # nobody may tear the exchanger down while a peer is still fetching the id
        x.put(f"ukernel_ccl/up/{rank}", True)
        x.count("ukernel_ccl/up/", world_size, 120000)
    finally:
        x.close()
    comm._exchanger = server  # rank 0 keeps it alive as long as the communicator
    return comm


def init_process_group(backend: str = "ukernel", comm: Optional[Communicator] = None, *, rank: Optional[int] = None,
                       world_size: Optional[int] = None, gpu_id: Optional[int] = None,
                       exchanger_ip: Optional[str] = None, exchanger_port: Optional[int] = None,
                       transport: str = "auto", heap_bytes: int = 1 << 30, stage_bytes: int = 64 << 20,
                       **kw) -> ProcessGroup:
    """Creates the default group.  Three ways to say which world it spans:

    * `comm`: an existing Communicator;
    * nothing, with ``torch.distributed`` initialised: ``Communicator.from_torch_dist()``;
    * the reference's keywords (ukernel_ccl/__init__.py:350-393) -- ``rank / world_size / gpu_id / exchanger_ip /
      exchanger_port`` with its defaults from ``RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT``: a
      stand-alone world rendezvoused through an exchanger that rank 0 serves.  `transport` is accepted ("auto";
      inside a box the transport is load/store over NVLink) and the reference's worker sizing knobs
      (``device_task_capacity, max_device_fifos, threads_per_block, fifo_capacity, smem_size``) are ignored.

    Other keyword arguments go to :class:`ProcessGroup` (``nlanes``, ``tile_bytes``, ``staging_bytes``)."""
    global _DEFAULT_GROUP
    if backend not in ("ukernel", "ucc", "ccl"):
        raise ValueError(f"unsupported backend {backend!r}")
    if _DEFAULT_GROUP is not None:
        raise RuntimeError("default ukernel process group already initialised")
    for knob in ("device_task_capacity", "max_device_fifos", "threads_per_block", "fifo_capacity", "smem_size"):
        kw.pop(knob, None)
    if comm is None:
        import os
        import torch.distributed as dist

        explicit = rank is not None or world_size is not None or exchanger_port is not None or exchanger_ip is not None
        if explicit or not (dist.is_available() and dist.is_initialized()):
            rank = int(os.environ.get("RANK", 0)) if rank is None else int(rank)
            world_size = int(os.environ.get("WORLD_SIZE", 1)) if world_size is None else int(world_size)
            gpu_id = int(os.environ.get("LOCAL_RANK", rank)) if gpu_id is None else int(gpu_id)
            exchanger_ip = os.environ.get("MASTER_ADDR", "127.0.0.1") if exchanger_ip is None else exchanger_ip
            exchanger_port = int(os.environ.get("MASTER_PORT", 29500)) if exchanger_port is None else int(exchanger_port)
            comm = _comm_from_exchanger(rank, world_size, gpu_id, exchanger_ip, exchanger_port, heap_bytes, stage_bytes)
        else:
            comm = Communicator.from_torch_dist(heap_bytes=heap_bytes, stage_bytes=stage_bytes)
    _DEFAULT_GROUP = ProcessGroup(comm, **kw)
    return _DEFAULT_GROUP


def _group(group: Optional[ProcessGroup]) -> ProcessGroup:
    pg = _DEFAULT_GROUP if group is None else group
    if pg is None:
        raise RuntimeError("ukernel process group is not initialised")
    return pg


def destroy_process_group(group: Optional[ProcessGroup] = None) -> None:
    global _DEFAULT_GROUP
    pg = _DEFAULT_GROUP if group is None else group
    if pg is not None:
        pg.shutdown()
    if group is None or group is _DEFAULT_GROUP:
        _DEFAULT_GROUP = None


def is_initialized() -> bool:
    return _DEFAULT_GROUP is not None


def get_rank(group: Optional[ProcessGroup] = None) -> int:
    return _group(group).rank


def get_world_size(group: Optional[ProcessGroup] = None) -> int:
    return _group(group).world_size


def barrier(group: Optional[ProcessGroup] = None, async_op: bool = False):
    return _group(group).barrier(async_op=async_op)


def all_reduce(tensor, op=ReduceOp.SUM, group: Optional[ProcessGroup] = None, async_op: bool = False, *,
               tile_bytes: Optional[int] = None, num_flows: Optional[int] = None):
    return _group(group).all_reduce(tensor, op=op, async_op=async_op, tile_bytes=tile_bytes, num_flows=num_flows)


def all_to_all_single(output, input, output_split_sizes=None, input_split_sizes=None,
                      group: Optional[ProcessGroup] = None, async_op: bool = False, *,
                      tile_bytes: Optional[int] = None, num_flows: Optional[int] = None):
    return _group(group).all_to_all_single(output, input, output_split_sizes, input_split_sizes, async_op=async_op,
                                           tile_bytes=tile_bytes, num_flows=num_flows)


class UkNetCommunicator:
    """ukernel collectives for ranks on *different boxes*: the same planner (ring / full-mesh tile DAGs), executed
    over the multipath datagram transport instead of load/store -- the role of the reference's TCP / UCCL
    transport adapters (experimental/ukernel/src/transport/adapter).  Tensors are host (or pinned) memory.

    ``exchange`` is an all-gather of small python objects between the members (as for ``net.NetCommunicator``);
    the communicator opens its own flows on ``engine``.
    """

    def __init__(self, rank: int, world_size: int, exchange, engine=None, nlanes: int = 2, tile_bytes: int = 1 << 20,
                 timeout_ms: int = 60000):
        from .. import net

        self._boot = net.NetCommunicator(rank, world_size, exchange, engine=engine or net.Engine(), timeout_ms=timeout_ms)
        flows = [self._boot.flows.get(p, 0) for p in range(world_size)]
        self.rank, self.world_size = rank, world_size
        self._u = _uk().UkNetComm(rank, world_size, self._boot.engine._native, flows, nlanes, int(tile_bytes), timeout_ms)

    @staticmethod
    def _host(t: torch.Tensor, name: str) -> torch.Tensor:
        if t.is_cuda or not t.is_contiguous():
            raise ValueError(f"ukernel net: {name} must be a contiguous host tensor")
        return t

    def all_reduce(self, tensor: torch.Tensor, op: str = "sum", out: Optional[torch.Tensor] = None, algo: str = "auto"):
        out = tensor if out is None else out
        self._u.all_reduce(self._host(tensor, "tensor").data_ptr(), self._host(out, "out").data_ptr(), tensor.numel(),
                           dtype_code(tensor.dtype), op_code(op), ALGOS[algo])
        return out

    def all_to_all_single(self, out: torch.Tensor, inp: torch.Tensor):
        self._u.all_to_all(self._host(inp, "inp").data_ptr(), self._host(out, "out").data_ptr(),
                           inp.numel() // self.world_size, dtype_code(inp.dtype))
        return out

    def all_gather_into_tensor(self, out: torch.Tensor, inp: torch.Tensor):
        self._u.all_gather(self._host(inp, "inp").data_ptr(), self._host(out, "out").data_ptr(), inp.numel(),
                           dtype_code(inp.dtype))
        return out

    def reduce_scatter_tensor(self, out: torch.Tensor, inp: torch.Tensor, op: str = "sum"):
        self._u.reduce_scatter(self._host(inp, "inp").data_ptr(), self._host(out, "out").data_ptr(), out.numel(),
                               dtype_code(inp.dtype), op_code(op))
        return out

    def broadcast(self, tensor: torch.Tensor, root: int = 0):
        self._u.broadcast(self._host(tensor, "tensor").data_ptr(), tensor.data_ptr(), tensor.numel(),
                          dtype_code(tensor.dtype), root)
        return tensor

    def barrier(self):
        self._u.barrier()

    def stats(self):
        return self._u.stats()
