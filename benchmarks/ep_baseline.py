#!/usr/bin/env python
"""EP dispatch/combine *baseline*: what a framework does without a fused EP library -- torch index ops to
pack tokens by destination rank, `dist.all_to_all_single` over NCCL (with the count exchange + host
sync that variable splits need), and an index_add to reduce on the way back.  Role of the reference's
ep/bench/baseline/* (torch.distributed / pack-unpack baselines).  Same shapes as bench.py, so the
numbers sit next to `uccl_b200.ep.Buffer`'s.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/ep_baseline.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tokens", type=int, default=4096)
    p.add_argument("--hidden", type=int, default=7168)
    p.add_argument("--topk", type=int, default=8)
    p.add_argument("--experts", type=int, default=256)
    p.add_argument("--iters", type=int, default=10)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, H, K, E, R = a.tokens, a.hidden, a.topk, a.experts, world
    g = torch.Generator(device=dev).manual_seed(rank)
    x = torch.randn(T, H, device=dev, generator=g).to(torch.bfloat16)
    idx = torch.rand(T, E, device=dev, generator=g).topk(K, dim=1).indices
    e_per = E // R

    def a2a(out, inp, out_splits, in_splits):
        if world > 1:
            dist.all_to_all_single(out, inp, out_splits, in_splits)
        else:
            out.copy_(inp)

    def dispatch():
        in_rank = torch.zeros(T, R, dtype=torch.bool, device=dev)
        in_rank.scatter_(1, idx // e_per, True)
        tok, dst = in_rank.nonzero(as_tuple=True)          # (token, rank) pairs
        order = torch.argsort(dst, stable=True)
        tok, dst = tok[order], dst[order]
        send_counts = torch.bincount(dst, minlength=R)
        recv_counts = torch.empty_like(send_counts)
        a2a(recv_counts, send_counts, None, None)
        sc, rc = send_counts.tolist(), recv_counts.tolist()  # host sync: NCCL needs the splits
        send = x.index_select(0, tok)
        recv = torch.empty(sum(rc), H, dtype=x.dtype, device=dev)
        a2a(recv, send, rc, sc)
        return recv, (tok, sc, rc)

    def combine(y, handle):
        tok, sc, rc = handle
        back = torch.empty(sum(sc), H, dtype=y.dtype, device=dev)
        a2a(back, y, sc, rc)
        out = torch.zeros(T, H, dtype=torch.float32, device=dev)
        out.index_add_(0, tok, back.float())
        return out.to(torch.bfloat16)

    def timed(fn, iters):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / iters * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    recv, h = dispatch()
    d_us = timed(dispatch, a.iters)
    c_us = timed(lambda: combine(recv, h), a.iters)
    res = {"impl": "torch+nccl all_to_all baseline", "n_gpus": world, "tokens": T, "hidden": H, "topk": K, "experts": E,
           "dispatch_us": d_us, "combine_us": c_us, "tokens_per_s": T * world / ((d_us + c_us) * 1e-6)}
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            json.dump(res, open(a.out, "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
