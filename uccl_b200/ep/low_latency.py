"""Low-latency (decode) dispatch / combine -- DeepEP ``low_latency_*`` API.

Reference: ep/bench/buffer.py:263-566 (python), ep/src/internode_ll.cu (kernels).
Return conventions follow DeepEP:
  dispatch -> (recv_x | (recv_x_fp8, scales), recv_count, handle, event, hook)
     recv_x  [E_local, R*M, H]; tokens of expert e are rows [0, recv_count[e]) (packed)
     scales  [E_local, R*M, H/128] float32 (row-major here; DeepEP hands out a transposed view)
     handle  (src_info, layout_range, M, H, E, buffer_idx, send_pos)
  combine  -> (combined_x [T, H] bf16, event, hook)
Only two LL buffers exist (as in DeepEP): at most two dispatch results may be alive.
"""
from __future__ import annotations

from typing import Optional

import torch

from .utils import EventHandle, EventOverlap


_LOGFMT_WARNED = False


def _warn_logfmt_once() -> None:
    """``use_logfmt=True`` is accepted for DeepEP API compatibility: the combine payload stays bf16 (a superset of
    LogFMT-10 in precision); over NVLink the 10-bit encoding would cost more SM time than the bytes it saves."""
    global _LOGFMT_WARNED
    if not _LOGFMT_WARNED:
        _LOGFMT_WARNED = True
        import warnings

        warnings.warn("uccl_b200.ep: use_logfmt=True is a no-op (payload stays bf16)", stacklevel=3)



def ll_size_hint(num_max_dispatch_tokens_per_rank: int, hidden: int, num_ranks: int, num_experts: int) -> int:
    from .. import _native

    return int(_native.C().EpBuffer.ll_size_hint(num_max_dispatch_tokens_per_rank, hidden, num_ranks, num_experts))


class LowLatencyRuntime:
    def __init__(self, buf, num_bytes: int):
        self.buf = buf
        self.rt = buf.runtime
        self.num_bytes = int(num_bytes)
        if self.num_bytes > 0:
            self.rt.ll_init(self.num_bytes)
        self._state = {}

    def _ensure(self, M, H, E):
        need = ll_size_hint(M, H, self.buf.group_size, E)
        if self.num_bytes == 0:
            self.num_bytes = need
            self.rt.ll_init(need)
        assert need <= self.num_bytes, f"low-latency buffer too small: need {need} bytes, have {self.num_bytes}"

    def clean(self, M, H, E):
        # signalling is epoch based: nothing to zero (reference needs clean_low_latency_buffer, buffer.py:1797)
        self._ensure(M, H, E)

    def _sms(self, combine: bool = False):
        from .buffer import Config

        # dispatch: one warp per token (8 warps / CTA); combine: one CTA per token
        return self.buf._sms(Config(128 if combine else min(self.buf.num_sms * 2, 64)))

    def dispatch(self, x: torch.Tensor, topk_idx: torch.Tensor, num_max_dispatch_tokens_per_rank: int,
                 num_experts: int, cumulative_local_expert_recv_stats: Optional[torch.Tensor] = None,
                 dispatch_wait_recv_cost_stats: Optional[torch.Tensor] = None, use_fp8: bool = True,
                 round_scale: bool = False, use_ue8m0: bool = False, async_finish: bool = False,
                 return_recv_hook: bool = False):
        if use_ue8m0:
            assert use_fp8 and round_scale, "use_ue8m0 needs use_fp8=True and round_scale=True (power-of-two scales)"
            assert x.size(1) % 512 == 0, "use_ue8m0 packs four per-128-channel scales per word: hidden % 512 == 0"
        assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.is_contiguous()
        assert topk_idx.dtype == torch.int64 and topk_idx.is_contiguous()
        b = self.buf
        R = b.group_size
        T, H = x.shape
        K = topk_idx.size(1)
        M, E = int(num_max_dispatch_tokens_per_rank), int(num_experts)
        E_local = E // R
        self._ensure(M, H, E)
        dev = b.device
        compute = b._enter(None, False)
        with torch.cuda.stream(b.comm_stream):
            recv_count = torch.empty(E_local, dtype=torch.int32, device=dev)
            layout_range = torch.empty((E_local, R), dtype=torch.int64, device=dev)
            send_pos = torch.empty((T, K), dtype=torch.int64, device=dev)
            rx, rs, rsrc, cx, idx = self.rt.ll_dispatch(x.data_ptr(), topk_idx.data_ptr(), T, H, K, E, M, use_fp8,
                                                        round_scale, recv_count.data_ptr(), layout_range.data_ptr(),
                                                        send_pos.data_ptr(), self._sms(), b.comm_stream.cuda_stream)
            if cumulative_local_expert_recv_stats is not None:
                cumulative_local_expert_recv_stats.add_(recv_count)
        rows = R * M
        if use_fp8:
            scales = b._view(rs, (E_local, rows, H // 128), torch.float32)
            if use_ue8m0:
                from .utils import pack_ue8m0

                with torch.cuda.stream(b.comm_stream):
                    scales = pack_ue8m0(scales)  # [E_local, rows, H // 512] int32, column-major last two dims
            recv_x = (b._view(rx, (E_local, rows, H), torch.float8_e4m3fn), scales)
        else:
            recv_x = b._view(rx, (E_local, rows, H), torch.bfloat16)
        src_info = b._view(rsrc, (E_local, rows), torch.int32)
        handle = (src_info, layout_range, M, H, E, idx, send_pos)
        ev = b._exit(compute, async_finish, (x, topk_idx, recv_count, layout_range, send_pos))
        hook = (lambda: None) if return_recv_hook else None
        return recv_x, recv_count, handle, ev, hook

    def next_combine_buffer(self, handle):
        src_info, layout_range, M, H, E, idx, send_pos = handle
        b = self.buf
        ptr = self.rt.ll_combine_buffer(idx, H, E, M)
        return b._view(ptr, (E // b.group_size, b.group_size * M, H), torch.bfloat16)

    def combine(self, x: torch.Tensor, topk_idx: torch.Tensor, topk_weights: torch.Tensor, handle,
                use_logfmt: bool = False, zero_copy: bool = False, async_finish: bool = False,
                return_recv_hook: bool = False, out: Optional[torch.Tensor] = None,
                combine_wait_recv_cost_stats: Optional[torch.Tensor] = None):
        if use_logfmt:
            _warn_logfmt_once()
        src_info, layout_range, M, H, E, idx, send_pos = handle
        b = self.buf
        assert x.dtype == torch.bfloat16 and x.is_contiguous() and x.shape[-1] == H
        assert topk_weights.dtype == torch.float32 and topk_weights.is_contiguous()
        T, K = topk_weights.shape
        compute = b._enter(None, False)
        with torch.cuda.stream(b.comm_stream):
            if out is None:
                out = torch.empty((T, H), dtype=torch.bfloat16, device=b.device)
            self.rt.ll_combine(x.data_ptr(), idx, topk_weights.data_ptr(), send_pos.data_ptr(), out.data_ptr(), T, H,
                               K, E, M, self._sms(combine=True), b.comm_stream.cuda_stream)
        ev = b._exit(compute, async_finish, (x, topk_weights, send_pos, out))
        hook = (lambda: None) if return_recv_hook else None
        return out, ev, hook
