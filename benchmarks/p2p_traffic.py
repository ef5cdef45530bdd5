#!/usr/bin/env python
"""Traffic-pattern generators on top of the P2P engine: **permutation** (every GPU sends to exactly one
other GPU) and **incast** (all GPUs send to GPU 0) -- the role of the reference's fabric experiments
(collective/rdma/azure_perm_traffic/permutation_traffic.cc, collective/rdma/incast/incast.cc), here for
the NVSwitch fabric: they show how one-sided writes of many endpoints share the switch.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/p2p_traffic.py --pattern permutation
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200.p2p import Endpoint


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pattern", default="permutation", choices=["permutation", "incast"])
    ap.add_argument("--bytes", type=int, default=64 << 20)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--shift", type=int, default=1, help="permutation: rank r sends to (r + shift) % N")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # out-of-band only (metadata + barriers)
    ep = Endpoint(local)
    dev = f"cuda:{local}"
    src = torch.full((a.bytes,), rank + 1, dtype=torch.uint8, device=dev)
    # one receive window per possible sender so that incast writers never overlap
    wins = [torch.zeros(a.bytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    descs = ep.register_memory(wins)
    mds = [None] * world
    dist.all_gather_object(mds, (ep.get_metadata(), ep.get_serialized_descs(descs)))
    if a.pattern == "permutation":
        dst = (rank + a.shift) % world
        sends = dst != rank
    else:
        dst = 0
        sends = rank != 0
    conn = None
    if sends:
        ok, conn = ep.connect(remote_metadata=mds[dst][0])
        assert ok
        remote = ep.deserialize_descs(mds[dst][1])[rank]
    if a.pattern == "permutation":
        n_in = 1 if (world > 1 and a.shift % world != 0) else 0
    else:
        n_in = world - 1 if rank == 0 else 0
    for _ in range(n_in):
        ep.accept(60000)
    local_desc = ep.register_memory([src])[0]
    torch.cuda.synchronize()
    dist.barrier()
    times = []
    for it in range(a.iters + 3):
        dist.barrier()
        t0 = time.perf_counter()
        if sends:
            ok, tid = ep.transfer(conn, "write", [local_desc], [remote])
            assert ok and ep.wait(tid, 60000)
        dt = time.perf_counter() - t0
        if it >= 3:
            times.append(dt)
    dist.barrier()
    torch.cuda.synchronize()
    ok_data = True
    if a.pattern == "permutation" and world > 1:
        s = (rank - a.shift) % world
        ok_data = bool((wins[s] == s + 1).all())
    elif a.pattern == "incast" and rank == 0:
        ok_data = all(bool((wins[s] == s + 1).all()) for s in range(1, world))
    mine = {"rank": rank, "sends": sends, "GBps": (a.bytes / (sum(times) / len(times)) / 1e9) if sends else 0.0,
            "data_ok": ok_data}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    if rank == 0:
        senders = [r for r in allr if r["sends"]]
        agg = sum(r["GBps"] for r in senders)
        res = {"pattern": a.pattern, "n_gpus": world, "bytes": a.bytes, "per_sender_GBps": [round(r["GBps"], 1) for r in senders],
               "aggregate_GBps": agg, "data_ok": all(r["data_ok"] for r in allr)}
        print(json.dumps(res))
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
