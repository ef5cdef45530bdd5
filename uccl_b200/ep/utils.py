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


def per_token_cast_back(x_fp8: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    m, n = x_fp8.shape
    xv = x_fp8.float().view(m, -1, 128)
    return (xv * scales.float().view(m, -1, 1)).view(m, n).to(torch.bfloat16)


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


def bench(fn, num_warmups: int = 5, num_tests: int = 20, flush_l2: bool = True):
    """Device-timed (CUDA events) avg/min/max seconds with an L2 flush between iterations."""
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
