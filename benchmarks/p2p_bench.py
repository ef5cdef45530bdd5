#!/usr/bin/env python
"""P2P KV-cache transfer sweep between two GPUs (BASELINE config #4: 128 KB - 1 GB), the counterpart of the
reference's p2p/benchmarks/benchmark_uccl.py --write-ipc / --read-ipc [--num-kvblocks N] [--async-api] [dual]
(p2p/benchmarks/benchmark_uccl.py:776-860).  Single process, two endpoints (GPU0 <-> GPU1).

Modes per (total bytes, num_kvblocks):
  write / read      one transfer() of the whole block vector, then wait()          (ONE kernel launch)
  prepared_write    prepare_transfer() once, then post_transfer() + wait() per iteration (NIXL prepXfer/postXfer)
  async             `--inflight` transfers issued back to back on the engine's side streams, then all waited
  dual              both endpoints write to each other at the same time (bidirectional NVLink load)
  memcpy            the copy-engine baseline the reference's engine uses: one cudaMemcpyAsync per block
Timing: host wall clock around issue + completion like the reference (benchmark_uccl.py:555-618), `--iters`
repetitions after 3 warm-ups; the memcpy baseline is timed the same way so launch overheads are compared too.

  python benchmarks/p2p_bench.py [--num-kvblocks 1,64,1024] [--out file.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uccl_b200.p2p import Endpoint


def timed(fn, iters, sync):
    for _ in range(3):
        fn()
    sync()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    sync()
    return (time.perf_counter() - t0) / iters


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--num-kvblocks", default="1,64,1024")
    p.add_argument("--sizes", default="131072,1048576,8388608,67108864,268435456,1073741824",
                   help="total bytes per transfer (split evenly over the kv blocks)")
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--inflight", type=int, default=4)
    p.add_argument("--out", default=None)
    args = p.parse_args()
    ng = torch.cuda.device_count()
    d0, d1 = 0, (1 if ng > 1 else 0)
    a, b = Endpoint(d0), Endpoint(d1)
    ok, conn_ab = a.connect(remote_metadata=b.get_metadata())
    assert ok
    b.accept(5000)
    ok, conn_ba = b.connect(remote_metadata=a.get_metadata())
    assert ok
    a.accept(5000)

    def sync():
        torch.cuda.synchronize(d0)
        torch.cuda.synchronize(d1)

    rows = []
    for nb in [int(v) for v in args.num_kvblocks.split(",")]:
        for size in [int(v) for v in args.sizes.split(",")]:
            blk = size // nb
            if blk < 4096:
                continue
            blk = blk // 16 * 16
            # one flat allocation per side, blocks are views (KV-cache pages of one pool)
            src0 = torch.ones(nb * blk, dtype=torch.uint8, device=f"cuda:{d0}")
            dst1 = torch.zeros(nb * blk, dtype=torch.uint8, device=f"cuda:{d1}")
            src1 = torch.ones(nb * blk, dtype=torch.uint8, device=f"cuda:{d1}")
            dst0 = torch.zeros(nb * blk, dtype=torch.uint8, device=f"cuda:{d0}")
            v = lambda t: [t[i * blk:(i + 1) * blk] for i in range(nb)]  # noqa: E731
            la = a.register_memory(v(src0))
            ra = a.deserialize_descs(b.get_serialized_descs(b.register_memory(v(dst1))))
            lb = b.register_memory(v(src1))
            rb = b.deserialize_descs(a.get_serialized_descs(a.register_memory(v(dst0))))
            sync()
            total = blk * nb
            iters = args.iters if total <= (64 << 20) else max(3, args.iters // 2)
            res = {"bytes": total, "blocks": nb, "block_bytes": blk}

            def one(op):
                ok_, tid = a.transfer(conn_ab, op, la, ra)
                a.wait(tid)

            for op in ("write", "read"):
                dt = timed(lambda: one(op), iters, sync)
                res[op] = {"us": dt * 1e6, "GBps": total / dt / 1e9}

            # prepared: descriptor lists resolved once, every post is one bare kernel launch
            prep = a.prepare_transfer(conn_ab, "write", la, ra)

            def posted():
                ok_, tid = a.post_transfer(prep)
                a.wait(tid)

            dt = timed(posted, iters, sync)
            res["prepared_write"] = {"us": dt * 1e6, "GBps": total / dt / 1e9}
            a.release_transfer(prep)

            def many():
                tids = [a.transfer(conn_ab, "write", la, ra)[1] for _ in range(args.inflight)]
                for t_ in tids:
                    a.wait(t_)

            dt = timed(many, iters, sync) / args.inflight
            res["async"] = {"us": dt * 1e6, "GBps": total / dt / 1e9, "inflight": args.inflight}

            def dual():
                _, t1 = a.transfer(conn_ab, "write", la, ra)
                _, t2 = b.transfer(conn_ba, "write", lb, rb)
                a.wait(t1)
                b.wait(t2)

            dt = timed(dual, iters, sync)
            res["dual"] = {"us": dt * 1e6, "GBps_per_direction": total / dt / 1e9}

            # copy-engine baseline: one cudaMemcpyAsync (peer copy) per block, like the reference's write_ipc
            s_blocks, d_blocks = v(src0), v(dst1)

            def memcpy():
                with torch.cuda.device(d0):
                    for s_, d_ in zip(s_blocks, d_blocks):
                        d_.copy_(s_, non_blocking=True)
                    torch.cuda.current_stream().synchronize()

            dt = timed(memcpy, iters, sync)
            res["memcpy_per_block"] = {"us": dt * 1e6, "GBps": total / dt / 1e9}
            res["speedup_write_vs_memcpy"] = res["memcpy_per_block"]["us"] / res["write"]["us"]
            assert bool((dst1 == 1).all()), "payload mismatch"
            rows.append(res)
            res["speedup_prepared_vs_memcpy"] = res["memcpy_per_block"]["us"] / res["prepared_write"]["us"]
            print(f"{total:>11d} B in {nb:4d} blocks: write {res['write']['GBps']:7.1f} GB/s ({res['write']['us']:8.1f} us) | "
                  f"prepared {res['prepared_write']['GBps']:7.1f} ({res['prepared_write']['us']:8.1f} us) | "
                  f"read {res['read']['GBps']:7.1f} | async {res['async']['GBps']:7.1f} | dual {res['dual']['GBps_per_direction']:7.1f}/dir | "
                  f"memcpy/block {res['memcpy_per_block']['GBps']:7.1f} GB/s ({res['memcpy_per_block']['us']:8.1f} us) "
                  f"-> x{res['speedup_write_vs_memcpy']:.2f}", flush=True)
            for ep, h in ((a, la), (b, lb)):
                ep.deregister_memory(h)
            del src0, dst1, src1, dst0
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"devices": [d0, d1], "rows": rows, "stats": a.stats()}, f, indent=1)


if __name__ == "__main__":
    main()
