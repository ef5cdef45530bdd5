#!/usr/bin/env python
"""Throughput and ratio of the lossless float codec (P2P compression hook) on one GPU.

    python benchmarks/compress_bench.py [--out f.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uccl_b200.p2p.compress import Compressor


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    comp = Compressor("for")
    rows = []
    for dtype in (torch.bfloat16, torch.float32):
        for numel in (1 << 20, 1 << 24, 1 << 27):
            x = torch.randn(numel, device=dev).to(dtype)
            buf = torch.empty(Compressor.bound(numel, dtype), dtype=torch.uint8, device=dev)
            out = torch.empty_like(x)
            _, nbytes = comp.compress(x, out=buf)

            def t(fn, iters=10):
                fn()
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(iters):
                    fn()
                e.record()
                torch.cuda.synchronize()
                return s.elapsed_time(e) / iters * 1e-3

            raw = numel * x.element_size()
            tc = t(lambda: comp.compress(x, out=buf))   # includes the size read-back (a stream sync)
            td = t(lambda: comp.decompress(buf, numel, dtype, out=out))
            assert torch.equal(out.view(torch.uint8), x.view(torch.uint8))
            row = {"dtype": str(dtype), "bytes": raw, "ratio": nbytes / raw, "compress_GBps": raw / tc / 1e9,
                   "decompress_GBps": raw / td / 1e9}
            rows.append(row)
            print(row)
    if a.out:
        json.dump({"rows": rows}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
