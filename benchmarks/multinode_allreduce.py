"""All-reduce across boxes: rail-aligned hierarchical algorithm (`parallel.MultiNodeCommunicator`) vs a flat
ring over the datagram transport, size sweep.

    # real cluster: one process per GPU, 8 per node
    torchrun --nnodes 2 --nproc-per-node 8 --rdzv-endpoint HOST:29500 benchmarks/multinode_allreduce.py
    # one box standing in for 2 boxes of 2 ranks (host backend when there is no GPU)
    torchrun --nproc-per-node 4 --master-addr 127.0.0.1 benchmarks/multinode_allreduce.py --local-size 2

The bootstrap world is gloo (any backend works; it only ships addresses).  Reported: algorithm bandwidth
(bytes / time) and bus bandwidth (x 2(n-1)/n) per size, max over ranks.  Reference role:
collective/rdma/run_nccl_test.sh (all_reduce_perf over the UCCL net plugin across hosts).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from uccl_b200 import net  # noqa: E402
from uccl_b200.parallel import MultiNodeCommunicator  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--local-size", type=int, default=int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
    ap.add_argument("--min-bytes", type=int, default=1 << 12)
    ap.add_argument("--max-bytes", type=int, default=1 << 26)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--bind", default="")
    ap.add_argument("--paths", type=int, default=0)
    ap.add_argument("--cc", default=None)
    ap.add_argument("--flat", action="store_true", help="also time a flat ring over the transport (every rank on the network)")
    args = ap.parse_args()
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.device("cuda") if cuda else torch.device("cpu")
    eng = net.Engine(bind_ip=args.bind, paths=args.paths, cc=args.cc)
    m = MultiNodeCommunicator.from_torch_dist(args.local_size, engine=eng, host=not cuda,
                                              heap_bytes=(2 << 30) if cuda else (256 << 20), stage_bytes=(64 << 20) if cuda else (8 << 20))
    flat = net.NetCommunicator.from_process_group(engine=net.Engine(bind_ip=args.bind, paths=args.paths, cc=args.cc)) if args.flat else None
    rows = []
    size = args.min_bytes
    while size <= args.max_bytes:
        x = torch.ones(size // 4, dtype=torch.float32, device=dev)
        res = {}
        for name, fn in (("hier", lambda: m.all_reduce(x)), ("flat", (lambda: flat.all_reduce(x)) if flat and not cuda else None)):
            if fn is None:
                continue
            for _ in range(2):
                fn()
            dist.barrier()
            t0 = time.perf_counter()
            for _ in range(args.iters):
                fn()
            if cuda:
                torch.cuda.synchronize()
            dt = torch.tensor([(time.perf_counter() - t0) / args.iters])
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            res[name] = float(dt)
        if rank == 0:
            row = {"bytes": size}
            for k, v in res.items():
                row[k + "_us"] = v * 1e6
                row[k + "_busbw_gbps"] = size * 8e-9 / v * 2 * (world - 1) / world
            rows.append(row)
            print(" ".join(f"{k}={v:.1f}" if isinstance(v, float) else f"{k}={v}" for k, v in row.items()), flush=True)
        size *= 4
    if rank == 0:
        print(json.dumps({"bench": "multinode_allreduce", "world": world, "local_size": args.local_size,
                          "nodes": world // args.local_size, "device": str(dev), "rows": rows}))
    m.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
