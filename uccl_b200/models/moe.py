"""Expert-parallel MoE layer on top of ``uccl_b200.ep.Buffer`` -- the consumer shape the
reference targets through Megatron / vLLM / SGLang (ep/bench/{megatron,vllm,sglang}).
Router -> get_dispatch_layout -> dispatch (optionally fused fp8) -> grouped expert MLP ->
zero-copy combine."""
from __future__ import annotations


import torch
import torch.nn as nn
import torch.nn.functional as F

from ..ep import Buffer, per_token_cast_back


class ExpertParallelMoE(nn.Module):
    def __init__(self, hidden: int, ffn: int, num_experts: int, top_k: int, buffer: Buffer,
                 dtype: torch.dtype = torch.bfloat16, use_fp8_dispatch: bool = False):
        super().__init__()
        R = buffer.group_size
        assert num_experts % R == 0
        self.hidden, self.ffn, self.num_experts, self.top_k = hidden, ffn, num_experts, top_k
        self.buffer = buffer
        self.local_experts = num_experts // R
        self.use_fp8_dispatch = use_fp8_dispatch
        self.router = nn.Linear(hidden, num_experts, bias=False, dtype=dtype)
        self.w1 = nn.Parameter(torch.randn(self.local_experts, hidden, ffn, dtype=dtype) * hidden ** -0.5)
        self.w2 = nn.Parameter(torch.randn(self.local_experts, ffn, hidden, dtype=dtype) * ffn ** -0.5)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """x: [tokens, hidden] bf16.  With gradients enabled the differentiable dispatch/combine of
        ``uccl_b200.ep.autograd`` is used (training); otherwise the zero-copy inference path."""
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._forward_train(x)
        with torch.no_grad():
            return self._forward_infer(x)

    def _forward_train(self, x: torch.Tensor) -> torch.Tensor:
        from ..ep.autograd import ep_combine, ep_dispatch

        logits = self.router(x).float()
        w, idx = torch.topk(F.softmax(logits, dim=-1), self.top_k, dim=-1)
        idx = idx.to(torch.int64).contiguous()
        recv_x, recv_idx, recv_w, _, handle = ep_dispatch(self.buffer, x.contiguous(), idx, w.float().contiguous(),
                                                          self.num_experts)
        out = torch.zeros(recv_x.size(0), self.hidden, dtype=recv_x.dtype, device=recv_x.device)
        for e in range(self.local_experts):
            sel = (recv_idx == e)
            rows = sel.any(dim=1).nonzero().flatten()
            if rows.numel() == 0:
                continue
            gate = (recv_w * sel).sum(dim=1)[rows].to(recv_x.dtype)
            h = F.silu(recv_x[rows] @ self.w1[e]) @ self.w2[e]
            out = out.index_add(0, rows, h * gate[:, None])
        return ep_combine(self.buffer, out, handle)

    def _forward_infer(self, x: torch.Tensor) -> torch.Tensor:
        buf = self.buffer
        logits = self.router(x).float()
        w, idx = torch.topk(F.softmax(logits, dim=-1), self.top_k, dim=-1)
        idx = idx.to(torch.int64).contiguous()
        w = w.float().contiguous()
        tpr, _, tpe, in_rank, _ = buf.get_dispatch_layout(idx, self.num_experts)
        recv_x, recv_idx, recv_w, per_expert, handle, _ = buf.dispatch(
            x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank, num_tokens_per_expert=tpe, topk_idx=idx,
            topk_weights=w, use_fp8=self.use_fp8_dispatch)
        if isinstance(recv_x, tuple):
            recv_x = per_token_cast_back(recv_x[0], recv_x[1])
        n = recv_x.size(0)
        out = buf.get_combine_buffer(n, self.hidden, self.top_k)
        out.zero_()
        # every received token is processed by each of its local experts, weighted by its gate
        for e in range(self.local_experts):
            sel = (recv_idx == e)
            rows = sel.any(dim=1).nonzero().flatten()
            if rows.numel() == 0:
                continue
            gate = (recv_w * sel).sum(dim=1)[rows].to(recv_x.dtype)
            h = F.silu(recv_x[rows] @ self.w1[e]) @ self.w2[e]
            out.index_add_(0, rows, h * gate[:, None])
        y, _, _ = buf.combine(out, handle)
        return y
