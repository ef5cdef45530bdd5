#!/usr/bin/env python
"""AllReduce on ORDINARY (cudaMalloc / torch) tensors: the staged kernels next to NCCL -- the default DDP path.
Compares the serial staged kernel (copy-in -> NVLS reduce -> copy-out per chunk) with the block-pipelined one
(three CTA groups run the phases concurrently) and checks the result.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/ar_plain_bench.py [--out f.json]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--sizes", default="33554432,67108864,134217728,268435456,1073741824")
    p.add_argument("--algos", default="auto,staged_nvls,staged_pipe")
    p.add_argument("--iters", type=int, default=8)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    rank, n, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    comm = Communicator.from_torch_dist(None, heap_bytes=1 << 30, stage_bytes=256 << 20)

    def mx(v):
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(a.iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return mx(s.elapsed_time(e) / a.iters) * 1e3

    rows = []
    for sz in [int(v) for v in a.sizes.split(",")]:
        x = torch.empty(sz // 2, dtype=torch.bfloat16, device=dev)
        row = {"bytes": sz}
        for algo in a.algos.split(","):
            try:
                x.fill_(float(rank + 1))
                comm.all_reduce(x, "sum", algo=algo)
                torch.cuda.synchronize()
                want = float(n * (n + 1) // 2)
                ok = bool((x[:4096] == want).all()) and bool((x[-4096:] == want).all()) and float(x[x.numel() // 3]) == want
                x.fill_(0.001)
                us = timed(lambda: comm.all_reduce(x, "sum", algo=algo))
                row[algo] = {"us": us, "busbw_GBps": sz / (us * 1e-6) * 2 * (n - 1) / n / 1e9, "ok": ok,
                             "picked": comm.select_allreduce(sz, False, torch.bfloat16)[0] if algo == "auto" else algo}
            except Exception as e:  # noqa: BLE001
                row[algo] = {"error": f"{type(e).__name__}: {e}"[:200]}
        x.fill_(0.001)
        us = timed(lambda: dist.all_reduce(x))
        row["nccl"] = {"us": us, "busbw_GBps": sz / (us * 1e-6) * 2 * (n - 1) / n / 1e9}
        rows.append(row)
        if rank == 0:
            print(json.dumps(row), flush=True)
        del x
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            json.dump({"n_gpus": n, "dtype": "bf16", "rows": rows}, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
