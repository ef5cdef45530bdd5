#!/usr/bin/env python
"""P2P KV-cache transfer sweep between two GPUs (BASELINE config #4: 128 KB - 1 GB blocks), the
counterpart of the reference's p2p/benchmarks/benchmark_uccl.py --write-ipc/--read-ipc
[--num-kvblocks N].  Single process, two endpoints (GPU0 -> GPU1); device-timed.

  python benchmarks/p2p_bench.py [--num-kvblocks 16] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uccl_b200.p2p import Endpoint


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--num-kvblocks", type=int, default=1)
    p.add_argument("--min", type=int, default=128 << 10)
    p.add_argument("--max", type=int, default=1 << 30)
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--out", default=None)
    args = p.parse_args()
    ng = torch.cuda.device_count()
    d0, d1 = 0, (1 if ng > 1 else 0)
    a, b = Endpoint(d0), Endpoint(d1)
    ok, conn = a.connect(remote_metadata=b.get_metadata())
    assert ok
    b.accept(5000)
    rows = []
    size = args.min
    while size <= args.max:
        nb = args.num_kvblocks
        blk = max(size // nb, 16)
        srcs = [torch.ones(blk, dtype=torch.uint8, device=f"cuda:{d0}") for _ in range(nb)]
        dsts = [torch.zeros(blk, dtype=torch.uint8, device=f"cuda:{d1}") for _ in range(nb)]
        local = a.register_memory(srcs)
        remote = a.deserialize_descs(b.get_serialized_descs(b.register_memory(dsts)))
        torch.cuda.synchronize(d0)
        torch.cuda.synchronize(d1)
        res = {"bytes": blk * nb, "blocks": nb}
        for op in ("write", "read"):
            for _ in range(2):
                ok, tid = a.transfer(conn, op, local, remote)
                a.wait(tid)
            t0 = time.perf_counter()
            for _ in range(args.iters):
                ok, tid = a.transfer(conn, op, local, remote)
                a.wait(tid)
            dt = (time.perf_counter() - t0) / args.iters
            res[op] = {"us": dt * 1e6, "GBps": blk * nb / dt / 1e9}
        # copy-engine baseline (what the reference's engine does: cudaMemcpyAsync peer copies)
        with torch.cuda.device(d0):
            for _ in range(2):
                for s_, d_ in zip(srcs, dsts):
                    d_.copy_(s_, non_blocking=True)
            torch.cuda.synchronize(d0)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(args.iters):
                for s_, d_ in zip(srcs, dsts):
                    d_.copy_(s_, non_blocking=True)
            e.record()
            torch.cuda.synchronize(d0)
            ms = s.elapsed_time(e) / args.iters
        res["memcpy_peer"] = {"us": ms * 1e3, "GBps": blk * nb / (ms * 1e-3) / 1e9}
        rows.append(res)
        print(f"{blk * nb:>11d} B x{nb}: write {res['write']['GBps']:7.1f} GB/s ({res['write']['us']:8.1f} us) | "
              f"read {res['read']['GBps']:7.1f} GB/s | cudaMemcpyPeer {res['memcpy_peer']['GBps']:7.1f} GB/s", flush=True)
        a.deregister_memory(local)
        size *= 4
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"devices": [d0, d1], "rows": rows, "stats": a.stats()}, f, indent=1)


if __name__ == "__main__":
    main()
