"""Lossless float compression hook for P2P transfers (``UCCL_P2P_COMPRESS_STRATEGY`` role of the
reference, p2p/rdma/compression.h).  Strategies: ``none`` (default) and ``for`` (exponent /
mantissa split + per-block frame-of-reference bit planes; bf16 and fp32).

    comp = Compressor()
    buf, nbytes = comp.compress(t)              # uint8 CUDA tensor + valid byte count
    out = comp.decompress(buf, t.numel(), t.dtype)

``Endpoint.send_compressed / recv_compressed`` (uccl_b200.p2p) use it for tensors of at least
``min_bytes`` when the strategy is not ``none``.
"""
from __future__ import annotations

import os
import struct
from typing import Tuple

import torch

from .. import _native
from ..parallel.comm import dtype_code

STRATEGIES = ("none", "for")


def default_strategy() -> str:
    s = os.environ.get("UCCL_B200_P2P_COMPRESS", os.environ.get("UCCL_P2P_COMPRESS_STRATEGY", "none")).lower()
    if s in ("split", "encode"):  # the reference's names map onto the one codec we ship
        s = "for"
    if s not in STRATEGIES:
        raise ValueError(f"unknown compression strategy {s!r} (known: {STRATEGIES})")
    return s


class Compressor:
    min_bytes = 2 << 20  # same threshold as the reference: small messages are never worth it

    def __init__(self, strategy: str | None = None):
        self.strategy = strategy or default_strategy()
        if self.strategy not in STRATEGIES:
            raise ValueError(f"unknown compression strategy {self.strategy!r}")

    @staticmethod
    def supports(t: torch.Tensor) -> bool:
        return t.is_cuda and t.is_contiguous() and t.dtype in (torch.bfloat16, torch.float32)

    def wants(self, t: torch.Tensor) -> bool:
        return self.strategy != "none" and self.supports(t) and t.numel() * t.element_size() >= self.min_bytes

    @staticmethod
    def bound(numel: int, dtype: torch.dtype) -> int:
        return _native.C().cmp_bound(int(numel), dtype_code(dtype))

    def compress(self, t: torch.Tensor, out: torch.Tensor | None = None) -> Tuple[torch.Tensor, int]:
        """Returns (buffer, nbytes).  Synchronises the current stream to read the size back."""
        if not self.supports(t):
            raise TypeError("Compressor: contiguous CUDA bf16/fp32 tensors only")
        C = _native.C()
        cap = self.bound(t.numel(), t.dtype)
        if out is None:
            out = torch.empty(cap, dtype=torch.uint8, device=t.device)
        assert out.dtype == torch.uint8 and out.numel() >= cap and out.is_contiguous()
        st = torch.cuda.current_stream(t.device)
        with torch.cuda.device(t.device):
            C.cmp_compress(t.data_ptr(), t.numel(), dtype_code(t.dtype), out.data_ptr(), st.cuda_stream)
            hdr = out[:C.CMP_HEADER_BYTES].cpu()  # stream-ordered D2H + sync
        total = struct.unpack_from("<Q", bytes(hdr.numpy().tobytes()), 24)[0]
        return out, int(total)

    def decompress(self, buf: torch.Tensor, numel: int, dtype: torch.dtype, out: torch.Tensor | None = None) -> torch.Tensor:
        C = _native.C()
        if out is None:
            out = torch.empty(numel, dtype=dtype, device=buf.device)
        assert out.is_contiguous() and out.numel() == numel and out.dtype == dtype
        with torch.cuda.device(buf.device):
            C.cmp_decompress(buf.data_ptr(), out.data_ptr(), int(numel), dtype_code(dtype),
                             torch.cuda.current_stream(buf.device).cuda_stream)
        return out
