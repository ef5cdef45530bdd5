#!/usr/bin/env python
"""Prefill → decode KV-cache hand-off over the P2P engine (the NIXL-style flow the reference's
`Endpoint.transfer` serves in vLLM/SGLang PD disaggregation, p2p/README.md).

Two processes on one node: the *decode* worker registers its paged KV blocks and publishes their
descriptors; the *prefill* worker fills its own blocks and pushes the ones the request needs with a
single vectorised one-sided write, then posts a notification.  Nothing is staged through host memory
and the decode side runs no code while the data moves.

    python examples/kv_transfer.py [--blocks 64] [--block-kb 256] [--prefill-gpu 0] [--decode-gpu 1]
"""
import argparse
import multiprocessing as mp
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def decode_worker(q_md, q_out, gpu, nblocks, block_bytes):
    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(gpu)
    ep = Endpoint(gpu)
    kv = [torch.zeros(block_bytes, dtype=torch.uint8, device=f"cuda:{gpu}") for _ in range(nblocks)]
    descs = ep.register_memory(kv)
    q_md.put((ep.get_metadata(), ep.get_serialized_descs(descs)))
    ok, ip, peer_gpu, conn = ep.accept(120000)
    assert ok
    done = None
    t0 = time.time()
    while done is None and time.time() - t0 < 120:  # the decode worker only polls for the notification
        for _, msg in ep.get_notifs():
            done = msg
        time.sleep(0.001)
    torch.cuda.synchronize()
    ids = [int(t) for t in done.decode().split(":")[1].split(",")]
    good = all(bool((kv[i] == (i % 251)).all()) for i in ids)
    untouched = all(bool((kv[i] == 0).all()) for i in range(nblocks) if i not in ids)
    q_out.put((good, untouched, len(ids)))


def prefill_worker(q_md, q_out, gpu, nblocks, block_bytes):
    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(gpu)
    ep = Endpoint(gpu)
    md, blob = q_md.get(timeout=120)
    ok, conn = ep.connect(remote_metadata=md)
    assert ok
    remote = ep.deserialize_descs(blob)
    kv = [torch.full((block_bytes,), i % 251, dtype=torch.uint8, device=f"cuda:{gpu}") for i in range(nblocks)]
    local = ep.register_memory(kv)
    torch.cuda.synchronize()
    ids = list(range(0, nblocks, 2))  # the request's block table: every other block
    t0 = time.perf_counter()
    ok, tid = ep.transfer(conn, "write", [local[i] for i in ids], [remote[i] for i in ids])
    assert ok and ep.wait(tid, 60000)
    dt = time.perf_counter() - t0
    ep.send_notif(conn, ("kv-ready:" + ",".join(map(str, ids))).encode())
    q_out.put((len(ids) * block_bytes / dt / 1e9, dt * 1e6))
    time.sleep(0.5)  # keep the endpoint alive until the decode side has verified


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--blocks", type=int, default=64)
    ap.add_argument("--block-kb", type=int, default=256)
    ap.add_argument("--prefill-gpu", type=int, default=0)
    ap.add_argument("--decode-gpu", type=int, default=1)
    a = ap.parse_args()
    import torch

    if torch.cuda.device_count() < 2:
        a.decode_gpu = a.prefill_gpu  # two processes on one GPU still go through CUDA IPC
    ctx = mp.get_context("spawn")
    q_md, q_d, q_p = ctx.Queue(), ctx.Queue(), ctx.Queue()
    bb = a.block_kb << 10
    ps = [ctx.Process(target=decode_worker, args=(q_md, q_d, a.decode_gpu, a.blocks, bb)),
          ctx.Process(target=prefill_worker, args=(q_md, q_p, a.prefill_gpu, a.blocks, bb))]
    [p.start() for p in ps]
    gbps, us = q_p.get(timeout=300)
    good, untouched, n = q_d.get(timeout=300)
    [p.join(30) for p in ps]
    print(f"pushed {n} KV blocks of {a.block_kb} KiB in {us:.0f} us ({gbps:.1f} GB/s); "
          f"payload verified={good}, other blocks untouched={untouched}")
    return 0 if good and untouched else 1


if __name__ == "__main__":
    sys.exit(main())
