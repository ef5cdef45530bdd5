#!/usr/bin/env python
"""Device-side timeline of one low-latency dispatch + combine (block 0 of this rank): where the microseconds go
(kernel begin -> count-exchange barrier -> sends -> arrival barrier ...).  Uses the in-kernel trace events of
the communicator (Communicator.enable_trace), so it costs nothing when off.

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 benchmarks/ll_trace.py [--grid 16]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator
from uccl_b200.ep import Buffer


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--tokens", type=int, default=128)
    p.add_argument("--hidden", type=int, default=7168)
    p.add_argument("--topk", type=int, default=8)
    p.add_argument("--experts", type=int, default=256)
    p.add_argument("--num-sms", type=int, default=0, help="Buffer.num_sms override (dispatch grid = 2x, capped at 64)")
    p.add_argument("--out", default=None)
    a = p.parse_args()
    rank, n, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if n > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    M, H, K, E = a.tokens, a.hidden, a.topk, a.experts
    ll = Buffer.get_low_latency_rdma_size_hint(M, H, n, E)
    comm = Communicator.from_torch_dist(None, heap_bytes=ll + (512 << 20), stage_bytes=16 << 20) if n > 1 else \
        Communicator.local_world(1, devices=[local], heap_bytes=ll + (512 << 20))[0]
    if a.num_sms:
        Buffer.set_num_sms(a.num_sms)
    buf = Buffer(comm=comm, num_nvl_bytes=1 << 20, num_rdma_bytes=ll, low_latency_mode=True)
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    x = torch.randn(M, H, generator=g).to(torch.bfloat16).to(dev)
    idx = (torch.randn(M, E, generator=g).abs() + 1).topk(K, dim=-1).indices.to(torch.int64).contiguous().to(dev)
    w = torch.rand(M, K, generator=g).float().to(dev)
    for it in range(6):
        if it == 5:
            torch.cuda.synchronize()
            if n > 1:
                dist.barrier()
            comm.enable_trace(1 << 14)
        rx, cnt, h, _, _ = buf.low_latency_dispatch(x, idx, M, E, use_fp8=True)
        cb = buf.get_next_low_latency_combine_buffer(h)
        out, _, _ = buf.low_latency_combine(cb, idx, w, h)
    torch.cuda.synchronize()
    ev = comm.dump_trace()
    comm.disable_trace()
    if rank == 0:
        b0 = [e for e in ev if e["block"] == 0]
        t0 = b0[0]["t_ns"] if b0 else 0
        rows = [{"us": (e["t_ns"] - t0) / 1e3, "event": e["event"], "aux": e["aux"]} for e in b0]
        nblocks = len({e["block"] for e in ev})
        last = max(e["t_ns"] for e in ev) if ev else t0
        res = {"n_gpus": n, "blocks_seen": nblocks, "span_us_all_blocks": (last - t0) / 1e3, "block0": rows}
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
