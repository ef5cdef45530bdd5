"""Autograd wrappers of EP dispatch / combine (training path).

The two collectives are each other's adjoint: the gradient of ``dispatch`` w.r.t. the tokens is a
``combine`` of the received-token gradients (an unweighted sum over the ranks a token was sent to), and
the gradient of ``combine`` is a cached-handle ``dispatch`` of the output gradients.  The gate weights
travel with the tokens, so their gradient comes back through combine's weight reduction.  This is how
Megatron-style MoE training uses a DeepEP-compatible buffer (reference consumers:
ep/bench/megatron/*, thirdparty/Primus launchers).

    recv_x, recv_idx, recv_w, per_expert, handle = ep_dispatch(buf, x, topk_idx, topk_weights, num_experts)
    y = ep_combine(buf, expert_out, handle)
"""
from __future__ import annotations

from typing import List, Tuple

import torch


class _Dispatch(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, topk_weights, buf, topk_idx, num_experts, expert_alignment, box):
        tpr, _, tpe, in_rank, _ = buf.get_dispatch_layout(topk_idx, num_experts)
        recv_x, recv_idx, recv_w, per_expert, handle, _ = buf.dispatch(
            x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank, num_tokens_per_expert=tpe, topk_idx=topk_idx,
            topk_weights=topk_weights, expert_alignment=expert_alignment)
        ctx.buf, ctx.handle = buf, handle
        ctx.mark_non_differentiable(recv_idx)
        box["per_expert"], box["handle"] = per_expert, handle  # python objects leave through the side channel
        # results are views of a recycled receive arena on CUDA: give autograd its own copies
        return recv_x.clone(), recv_idx.clone(), recv_w.clone()

    @staticmethod
    def backward(ctx, g_x, _g_idx, g_w):
        buf, handle = ctx.buf, ctx.handle
        num_recv = handle.num_recv
        gx = g_x.contiguous() if g_x is not None else None
        if gx is None:
            return None, None, None, None, None, None, None
        gw = g_w.contiguous().float() if g_w is not None else None
        cin = buf.get_combine_buffer(num_recv, gx.size(1), handle.num_topk)
        cin[:num_recv].copy_(gx[:num_recv])
        grad_x, grad_w, _ = buf.combine(cin, handle, topk_weights=gw)
        return grad_x, grad_w, None, None, None, None, None


class _Combine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, expert_out, buf, handle):
        ctx.buf, ctx.handle = buf, handle
        num_recv = handle.num_recv
        cin = buf.get_combine_buffer(num_recv, expert_out.size(1), handle.num_topk)
        cin[:num_recv].copy_(expert_out[:num_recv])
        y, _, _ = buf.combine(cin, handle)
        return y

    @staticmethod
    def backward(ctx, g_y):
        buf, handle = ctx.buf, ctx.handle
        g, *_ = buf.dispatch(g_y.contiguous(), handle=handle)
        return g.clone(), None, None


def ep_dispatch(buf, x: torch.Tensor, topk_idx: torch.Tensor, topk_weights: torch.Tensor, num_experts: int,
                expert_alignment: int = 1) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, List[int], Tuple]:
    """Differentiable dispatch (bf16 payload).  Returns ``(recv_x, recv_topk_idx, recv_topk_weights,
    num_recv_tokens_per_expert_list, handle)``; gradients flow to ``x`` and ``topk_weights``."""
    assert x.dtype == torch.bfloat16 and topk_weights.dtype == torch.float32
    box = {}
    recv_x, recv_idx, recv_w = _Dispatch.apply(x, topk_weights, buf, topk_idx, num_experts, expert_alignment, box)
    return recv_x, recv_idx, recv_w, box["per_expert"], box["handle"]


def ep_combine(buf, expert_out: torch.Tensor, handle: Tuple) -> torch.Tensor:
    """Differentiable combine: ``y[t] = sum over the ranks token t was dispatched to of expert_out rows``."""
    assert expert_out.dtype == torch.bfloat16
    return _Combine.apply(expert_out, buf, handle)
