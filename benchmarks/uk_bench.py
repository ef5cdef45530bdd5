#!/usr/bin/env python
"""ukernel (persistent worker, no kernel launches on the caller's side) vs the kernel-launch collectives,
small to medium messages.  Host-timed end to end (enqueue -> result visible), because the point of the
ukernel path is what the *caller* pays.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/uk_bench.py [--out f.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator
from uccl_b200 import ukernel as uk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--lanes", type=int, default=4)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    comm = Communicator.from_torch_dist(heap_bytes=2 << 30, stage_bytes=64 << 20)
    u = uk.UkCommunicator(comm, nlanes=a.lanes, tile_bytes=256 << 10, staging_bytes=16 << 20)
    rows = []
    for nbytes in (4096, 65536, 1 << 20, 8 << 20):
        n = nbytes // 2
        x = comm.empty(n, dtype=torch.bfloat16)
        x.fill_(1)
        torch.cuda.synchronize()

        def timed(fn):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(a.iters):
                fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / a.iters * 1e6
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        row = {"bytes": nbytes,
               "kernel_us": timed(lambda: comm.all_reduce(x, "sum")),
               "ukernel_fullmesh_us": timed(lambda: u.all_reduce(x, "sum", symmetric=True).wait()),
               "ukernel_ring_us": timed(lambda: u.all_reduce(x, "sum", algo="ring", symmetric=True).wait()),
               "nccl_us": timed(lambda: dist.all_reduce(x))}
        rows.append(row)
        if rank == 0:
            print(row)
    res = {"n_gpus": world, "lanes": a.lanes, "rows": rows, "worker_launches": None}
    u.stop()
    if rank == 0 and a.out:
        json.dump(res, open(a.out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
