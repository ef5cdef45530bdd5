"""EP helpers with the semantics of the reference's ep/bench/utils.py (EventOverlap :~300,
per_token_cast_to_fp8 :666-675, calc_diff :36, bench :375-405), rewritten for this package."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


class EventHandle:
    """A CUDA event recorded on a stream (reference: ep/include/ep_event.hpp:8-45)."""

    def __init__(self, stream: Optional[torch.cuda.Stream] = None):
        self.event = torch.cuda.Event()
        self.event.record(stream if stream is not None else torch.cuda.current_stream())

    def current_stream_wait(self) -> None:
        torch.cuda.current_stream().wait_event(self.event)


class EpHandle(tuple):
    """Handle of an intranode dispatch, in the reference's field order (ep/bench/buffer.py:1147-1158):

        (rank_prefix_matrix, channel_prefix_matrix, recv_channel_prefix_matrix, num_recv_tokens,
         recv_src_idx, is_token_in_rank, send_head)

    so code that unpacks or indexes a DeepEP / uccl.ep handle keeps working.  This library has no channels
    (tokens are placed directly), so the two channel matrices are [R, 1] placeholders, and ``send_head``
    -- [num_tokens, num_ranks] int32 like DeepEP's -- holds the destination slot of every token at every
    rank (-1: not routed there).  ``slot`` (receive arena of the dispatch) and ``num_topk`` ride along as
    attributes."""

    def __new__(cls, rank_prefix, num_recv, recv_src_idx, is_token_in_rank, send_slot, slot=0, num_topk=0):
        R = rank_prefix.size(0) if hasattr(rank_prefix, "size") else 1
        import torch as _t

        dummy = _t.zeros((R, 1), dtype=_t.int32, device=getattr(rank_prefix, "device", None))
        self = super().__new__(cls, (rank_prefix, dummy, dummy, int(num_recv), recv_src_idx, is_token_in_rank, send_slot))
        self.slot = int(slot)
        self.num_topk = int(num_topk)
        return self

    rank_prefix = property(lambda self: self[0])
    num_recv = property(lambda self: self[3])
    recv_src_idx = property(lambda self: self[4])
    is_token_in_rank = property(lambda self: self[5])
    send_slot = property(lambda self: self[6])


class EventOverlap:
    """Returned by every Buffer call; lets the caller overlap communication with compute.

    ``with event_overlap: ...`` runs the body and *then* makes the current stream wait for the
    communication (DeepEP convention).  ``extra_tensors`` keeps buffers alive until waited on."""

    def __init__(self, event: Optional[EventHandle] = None, extra_tensors: Optional[Tuple] = None):
        self.event = event
        self.extra_tensors = extra_tensors

    def current_stream_wait(self) -> None:
        if self.event is not None:
            self.event.current_stream_wait()

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc, tb):
        if self.event is not None:
            self.event.current_stream_wait()
        return False


def per_token_cast_to_fp8(x: torch.Tensor, round_scale: bool = False):
    """[T, H] bf16/fp32 -> (e4m3 [T, H], inverse scales [T, H/128] fp32); amax clamped at 1e-4."""
    assert x.dim() == 2 and x.size(1) % 128 == 0
    m, n = x.shape
    xv = x.float().view(m, -1, 128)
    amax = xv.abs().amax(dim=2).clamp(1e-4)
    if round_scale:
        scale_inv = torch.pow(2.0, torch.ceil(torch.log2(amax / 448.0)))
        scale = 1.0 / scale_inv
    else:
        scale = 448.0 / amax
        scale_inv = amax / 448.0
    q = (xv * scale.unsqueeze(2)).to(torch.float8_e4m3fn).view(m, n)
    return q, scale_inv


def per_token_cast_back(x_fp8: torch.Tensor, x_scales: torch.Tensor = None, *, scales: torch.Tensor = None) -> torch.Tensor:
    """Inverse of ``per_token_cast_to_fp8`` (reference ep/bench/utils.py:per_token_cast_back); ``scales`` is the
    older spelling of ``x_scales``."""
    if x_scales is None:
        x_scales = scales
    if x_scales is None:
        raise TypeError("per_token_cast_back needs x_scales")
    m, n = x_fp8.shape
    xv = x_fp8.float().view(m, -1, 128)
    return (xv * x_scales.float().view(m, -1, 1)).view(m, n).to(torch.bfloat16)


def logfmt10_simulate(x: torch.Tensor) -> torch.Tensor:
    """The reference's LogFMT-10 "simulated cast" of the low-latency combine payload (ep/src/internode_ll.cu:934-995):
    per group of 128 channels with ``|max| <= 1`` every value is snapped to a 9-bit logarithmic grid between the
    group's smallest and largest magnitude (range clipped to ``2**-32`` of the maximum), the sign is kept and the
    result is bf16 again -- the wire format does not change, only the numerics.  Groups with ``|max| > 1`` (or a single
    magnitude) pass through.  This is the fp32 definition the CUDA kernel (``ep_ll_pack_logfmt_kernel``) is tested
    against; the host backend uses it directly."""
    assert x.dtype == torch.bfloat16 and x.shape[-1] % 128 == 0
    xf = x.float().reshape(*x.shape[:-1], x.shape[-1] // 128, 128)
    a = xf.abs()
    la = torch.log2(a)  # -inf at 0
    amax = a.amax(-1, keepdim=True)
    lmax = la.amax(-1, keepdim=True)
    inf = torch.full_like(la, float("inf"))
    lmin = torch.where(a > 0, la, inf).amin(-1, keepdim=True)
    lmin = torch.maximum(lmin, lmax - 32.0)
    step = (lmax - lmin) / 510.0
    step_inv = 1.0 / step
    rounding = 2.0 - torch.log2((1.0 + torch.exp2(step)) * 0.5) * step_inv
    enc = torch.floor((la - lmin) * step_inv + rounding)
    dec = torch.exp2((enc - 1.0) * step + lmin).to(torch.bfloat16).float()
    use = (amax <= 1.0) & (lmin < lmax)
    out = torch.where(use, torch.copysign(dec, xf), xf)
    return out.reshape(x.shape).to(torch.bfloat16)


def calc_diff(x: torch.Tensor, y: torch.Tensor) -> float:
    x, y = x.double() + 1, y.double() + 1
    denom = (x * x + y * y).sum()
    return float(1 - 2 * (x * y).sum() / denom)


def inplace_unique(x: torch.Tensor, num_slots: int) -> None:
    """Row-wise unique of non-negative ids, padded with -1 (used by tests to derive rank sets)."""
    assert x.dim() == 2
    mask = x < 0
    x_padded = x.masked_fill(mask, num_slots)
    bins = torch.zeros((x.size(0), num_slots + 1), dtype=x.dtype, device=x.device)
    bins.scatter_add_(1, x_padded, torch.ones_like(x_padded))
    bins = bins[:, :num_slots]
    sorted_bins, sorted_idx = torch.sort(bins > 0, dim=-1, descending=True, stable=True)
    sorted_idx = sorted_idx.masked_fill(~sorted_bins, -1)
    valid = min(num_slots, x.size(1))
    x[:, :] = -1
    x[:, :valid] = sorted_idx[:, :valid]


def bench(fn, num_warmups: int = 5, num_tests: int = 20, post_fn=None, flush_l2: bool = True):
    """Device-timed (CUDA events) avg/min/max seconds with an L2 flush between iterations; ``post_fn`` runs after
    each timed call, outside its event pair (reference ep/bench/utils.py:bench)."""
    torch.cuda.synchronize()
    cache = torch.empty(int(256e6 // 4), dtype=torch.int, device="cuda") if flush_l2 else None
    for _ in range(num_warmups):
        fn()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(num_tests)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(num_tests)]
    for i in range(num_tests):
        if cache is not None:
            cache.zero_()
        starts[i].record()
        fn()
        ends[i].record()
        if post_fn is not None:
            post_fn()
    torch.cuda.synchronize()
    ts = [s.elapsed_time(e) / 1e3 for s, e in zip(starts, ends)]
    return sum(ts) / len(ts), min(ts), max(ts)


def pack_ue8m0(scales: torch.Tensor) -> torch.Tensor:
    """Power-of-two fp32 scales ``[..., rows, C]`` (C % 4 == 0) -> UE8M0: the biased exponent byte of each
    scale, four per ``int32``, shaped ``[..., rows, C // 4]`` with the last two dimensions stored
    column-major (TMA-friendly; reference: ``low_latency_dispatch(use_ue8m0=True)``,
    ep/bench/buffer.py:301-316 -- the scale format SM100 block-scaled GEMMs consume)."""
    assert scales.dtype == torch.float32 and scales.size(-1) % 4 == 0
    bits = scales.contiguous().view(torch.int32)
    exp = ((bits >> 23) & 0xFF).to(torch.uint8)  # sign is 0 and the mantissa is 0 for power-of-two scales
    packed = exp.view(torch.int32)  # little endian: scale 4j+i sits in byte i of word j
    storage = packed.transpose(-1, -2).contiguous()
    return storage.transpose(-1, -2)


def unpack_ue8m0(packed: torch.Tensor) -> torch.Tensor:
    """Inverse of :func:`pack_ue8m0`: ``[..., rows, C // 4]`` int32 -> fp32 scales ``[..., rows, C]``."""
    assert packed.dtype == torch.int32
    exp = packed.contiguous().view(torch.uint8).to(torch.int32)
    return (exp << 23).view(torch.float32)


# ---------------------------------------------------------------------------------------------------------
# Harness helpers with the reference's names (ep/bench/utils.py) so that its test / benchmark scripts port over.
def hash_tensor(t: torch.Tensor) -> int:
    """Order-independent 64-bit checksum of a tensor's bytes (determinism checks across variants)."""
    b = t.contiguous().view(torch.uint8).to(torch.int64)
    pad = (-b.numel()) % 8
    if pad:
        b = torch.cat([b.reshape(-1), torch.zeros(pad, dtype=torch.int64, device=b.device)])
    w = b.reshape(-1, 8)
    shifts = torch.arange(8, device=b.device, dtype=torch.int64) * 8
    return int((w << shifts).sum().item() & 0xFFFFFFFFFFFFFFFF)


def create_grouped_scores(scores: torch.Tensor, group_idx: torch.Tensor, num_groups: int) -> torch.Tensor:
    """Keep only the scores of the selected expert groups (group-limited routing, DeepSeek-V3 style)."""
    num_tokens, num_experts = scores.shape
    s = scores.view(num_tokens, num_groups, -1)
    mask = torch.zeros((num_tokens, num_groups), dtype=torch.bool, device=scores.device)
    mask = mask.scatter_(1, group_idx, True).unsqueeze(-1).expand_as(s)
    return (s * mask).view(num_tokens, num_experts)


def init_dist(local_rank: int, num_local_ranks: int, backend: Optional[str] = None):
    """``(rank, world_size, group)`` from the usual MASTER_ADDR / MASTER_PORT / WORLD_SIZE / RANK variables
    (WORLD_SIZE = number of nodes and RANK = node index, as in the reference's launcher convention)."""
    import os

    import torch.distributed as dist

    ip = os.getenv("MASTER_ADDR", "127.0.0.1")
    port = int(os.getenv("MASTER_PORT", "8361"))
    num_nodes = int(os.getenv("WORLD_SIZE", 1))
    node_rank = int(os.getenv("RANK", 0))
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local_rank)
    dist.init_process_group(backend=backend or ("cpu:gloo,cuda:nccl" if cuda else "gloo"),
                            init_method=f"tcp://{ip}:{port}", world_size=num_nodes * num_local_ranks,
                            rank=node_rank * num_local_ranks + local_rank)
    return dist.get_rank(), dist.get_world_size(), dist.new_group(list(range(num_local_ranks * num_nodes)))


def init_dist_under_torchrun(local_rank: Optional[int] = None, num_local_ranks: Optional[int] = None):
    import os

    import torch.distributed as dist

    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", local_rank or 0)))
    dist.init_process_group("cpu:gloo,cuda:nccl" if cuda else "gloo")
    return dist.get_rank(), dist.get_world_size(), dist.new_group(list(range(dist.get_world_size())))


def detect_group_topology(group=None) -> Tuple[int, int, int, bool]:
    """``(num_nodes, ranks_per_node, node_index, is_intranode)``: one NVLink domain is one "node" here."""
    import torch.distributed as dist

    n = dist.get_world_size(group) if dist.is_initialized() else 1
    return 1, n, 0, True


def check_nvlink_connections(group=None) -> bool:
    """True when every pair of visible GPUs can reach each other over P2P (NVLink / NVSwitch)."""
    if not torch.cuda.is_available():
        return False
    n = torch.cuda.device_count()
    return all(torch.cuda.can_device_access_peer(a, b) for a in range(n) for b in range(n) if a != b)


def initialize_uccl(scratch_ptr=None, scratch_nbytes: int = 0, rank: int = 0, num_ranks: int = 1, group=None,
                    num_experts: int = 0, is_intranode=None, use_normal_mode: bool = False,
                    rdma_buffer_is_host_allocated: bool = False, **kw):
    """The reference spawns CPU proxy threads and exchanges their metadata here; kernels address peers directly in
    this library, so there is nothing to start.  Kept for script compatibility: returns ``([], None)``."""
    return [], None


def destroy_uccl(proxies=None, workers=None) -> None:
    return None


class suppress_stdout_stderr:
    """Silence both C-level and Python-level stdout / stderr inside the block (profiler chatter)."""

    def __enter__(self):
        import os
        import sys

        sys.stdout.flush()
        sys.stderr.flush()
        self._null = [os.open(os.devnull, os.O_RDWR) for _ in range(2)]
        self._saved = [os.dup(1), os.dup(2)]
        os.dup2(self._null[0], 1)
        os.dup2(self._null[1], 2)
        return self

    def __exit__(self, *exc):
        import os
        import sys

        sys.stdout.flush()
        sys.stderr.flush()
        os.dup2(self._saved[0], 1)
        os.dup2(self._saved[1], 2)
        for fd in self._null + self._saved:
            os.close(fd)
        return False


class empty_suppress:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


def bench_kineto(fn, kernel_names, num_tests: int = 30, suppress_kineto_output: bool = False,
                 trace_path: Optional[str] = None, barrier_comm_profiling: bool = False, num_kernels_per_period: int = 1):
    """Average device time (seconds) of the kernels whose names contain ``kernel_names`` over ``num_tests`` calls of
    ``fn``, measured with the torch profiler (splits dispatch from combine inside one call, like the reference's
    ep/bench/utils.py:408-544).  A tuple of names returns a tuple of times."""
    names = (kernel_names,) if isinstance(kernel_names, str) else tuple(kernel_names)
    ctx = suppress_stdout_stderr if suppress_kineto_output else empty_suppress
    with ctx():
        sched = torch.profiler.schedule(wait=0, warmup=1, active=1, repeat=1)
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA], schedule=sched) as prof:
            for _ in range(2):
                if barrier_comm_profiling:
                    import torch.distributed as dist

                    dev = torch.cuda.current_device()
                    lhs = torch.randn((4096, 4096), device=f"cuda:{dev}")
                    lhs @ lhs  # soak up launch skew before the measured region
                    if dist.is_initialized():
                        dist.all_reduce(torch.ones(1, device=f"cuda:{dev}"))
                for _ in range(num_tests):
                    fn()
                torch.cuda.synchronize()
                prof.step()
    if trace_path is not None:
        prof.export_chrome_trace(trace_path)
    events = prof.key_averages()
    out = []
    for name in names:
        tot, cnt = 0.0, 0
        for ev in events:
            if name in ev.key:
                t = getattr(ev, "device_time_total", None)
                if t is None:
                    t = getattr(ev, "cuda_time_total", 0.0)
                tot += float(t)
                cnt += int(ev.count)
        out.append(tot / max(cnt, 1) * 1e-6 * num_kernels_per_period)
    return out[0] if isinstance(kernel_names, str) else tuple(out)


# ---------------------------------------------------------------------------------------------------------------
# Bootstrap helpers of the reference's ep/bench/utils.py that launch scripts import (detect_ib_hca :293-320,
# get_peer_ip :139-148, get_cpu_proxies_meta :151-180).  Inside one NVSwitch box nothing of this is needed to build a
# Buffer; they exist so that those scripts run unchanged and for groups that span boxes.
def detect_ib_hca() -> Optional[str]:
    """Name of the first RDMA device of the host (Mellanox `mlx5_*` first, then Intel `irdma*`, then anything else),
    or None.  The scale-out transport of this library binds rails by network interface instead
    (`uccl_b200.net.topology.nic_for_gpu`), so this is informational."""
    import glob
    import os

    try:
        names = sorted(os.path.basename(p) for p in glob.glob("/sys/class/infiniband/*"))
    except OSError:
        return None
    for prefix in ("mlx5", "irdma", ""):
        for n in names:
            if n.startswith(prefix):
                return n
    return None


def _gather_objects(obj, num_ranks: int, group):
    if num_ranks <= 1:
        return [obj]
    import torch.distributed as dist

    out = [None] * num_ranks
    dist.all_gather_object(out, obj, group=group)
    return out


def get_peer_ip(rank: int, num_ranks: int, group=None) -> str:
    """Out-of-band address of the next rank of the ring ("" for a single rank)."""
    if num_ranks <= 1:
        return ""
    from ..p2p import get_oob_ip

    ips = _gather_objects(get_oob_ip(), num_ranks, group)
    return ips[(rank + 1) % num_ranks] or ""


def get_cpu_proxies_meta(proxies, rank: int, scratch_ptr: int, scratch_bytes: int, num_ranks: int, group=None) -> dict:
    """``{rank: {rank, ptr, nbytes, ip, listen_ports}}`` of every rank -- what the reference exchanges before it wires
    its proxies to each other (a :class:`uccl_b200.ep.Proxy` has no listen port: 0 is reported)."""
    from ..p2p import get_oob_ip

    meta = dict(rank=int(rank), ptr=int(scratch_ptr), nbytes=int(scratch_bytes), ip=get_oob_ip(),
                listen_ports=[int(getattr(p, "get_listen_port", lambda: 0)()) for p in (proxies or [])])
    return {m["rank"]: m for m in _gather_objects(meta, num_ranks, group)}
