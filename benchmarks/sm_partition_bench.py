#!/usr/bin/env python
"""SM partitions (CUDA green contexts, `uccl_b200.utils.SmPartition`): what a physical SM budget buys.

1. split table: what the driver grants for a requested SM count (granularity) and which SM ids a probe lands on;
2. copy bandwidth of a partition as a function of its size (how many SMs a copy-shaped kernel needs);
3. interference: an expert-parallel dispatch + combine at 24 SMs next to a GEMM stream, (a) both on ordinary streams,
   (b) the EP kernels on a 24-SM partition and the GEMMs on the rest.  Reported: EP step time alone / next to the GEMMs,
   GEMM time alone / next to the EP traffic.

    python benchmarks/sm_partition_bench.py                      # 1 GPU: parts 1 and 2, part 3 with one rank
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 benchmarks/sm_partition_bench.py --out f.json

Reference counterpart: experimental/misc/cuda_greenctx.cu (split + SM-id listing), cuda_concurrent.cu.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator
from uccl_b200.ep import Buffer, Config
from uccl_b200.utils import SmPartition, sm_ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sms", type=int, default=24, help="SM budget of the communication kernels")
    ap.add_argument("--tokens", type=int, default=4096)
    ap.add_argument("--hidden", type=int, default=7168)
    ap.add_argument("--topk", type=int, default=8)
    ap.add_argument("--experts", type=int, default=256)
    ap.add_argument("--gemm", type=int, default=8192, help="M = N = K of the competing bf16 GEMM")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rank, n, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if n > 1:
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
    ok, why = SmPartition.supported(local)
    if not ok:
        if rank == 0:
            print(json.dumps({"unavailable": why}))
        return 0
    res = {"n_gpus": n, "device": torch.cuda.get_device_name(local)}

    def mx(v):
        if n == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    # ---- 1. what a request turns into (each split is released before the next one)
    table = []
    for want in (8, 16, 24, 32, 64):
        for fine in (False, True):
            part, rest = SmPartition.split(want, local, fine_grained=fine)
            ids = part.sm_ids().tolist()
            table.append({"asked": want, "fine_grained": fine, "granted": part.sm_count, "rest": rest.sm_count if rest else 0,
                          "sm_ids": ids})
            del part, rest
    res["split"] = table
    res["default_stream_sms"] = int(sm_ids(blocks=8 * torch.cuda.get_device_properties(local).multi_processor_count).numel())

    # ---- 2. copy bandwidth vs partition size (256 MiB, larger than L2)
    src = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    dst = torch.empty_like(src)
    bw = []
    for want in (8, 16, 24, 32, 48, 64, 96):
        part, rest = SmPartition.split(want, local)
        st = part.stream()
        with torch.cuda.stream(st):
            for _ in range(3):
                dst.copy_(src)
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(10):
                dst.copy_(src)
            e.record()
        st.synchronize()
        bw.append({"sms": part.sm_count, "copy_GBps": 2 * src.numel() * 10 / (s.elapsed_time(e) * 1e-3) / 1e9})
        del part, rest
    res["copy_bw"] = bw

    # ---- 3. EP step next to a GEMM stream
    T, H, K, E = a.tokens, a.hidden, a.topk, a.experts
    arena = n * T * (H * 2 + K * 4) + (1 << 20)
    nvl = 3 * arena + (2 << 20)
    heap = nvl + (1 << 30)
    comm = Communicator.from_torch_dist(None, heap_bytes=heap, stage_bytes=64 << 20) if n > 1 else \
        Communicator.local_world(1, devices=[local], heap_bytes=heap)[0]
    buf = Buffer(comm=comm, num_nvl_bytes=nvl)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16).to(dev)
    scores = torch.randn(T, E, generator=g).abs() + 1
    idx = scores.topk(K, dim=-1).indices.to(torch.int64).contiguous().to(dev)
    w = torch.rand(T, K, generator=g).float().to(dev)
    part, rest = SmPartition.split(a.sms, local)
    cfg = Config(min(a.sms, part.sm_count))
    tpr, _, tpe, in_rank, _ = buf.get_dispatch_layout(idx, E)
    rx, _, _, _, handle, _ = buf.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank, num_tokens_per_expert=tpe,
                                          topk_idx=idx, topk_weights=w, config=cfg)
    cin = buf.get_combine_buffer(handle.num_recv, H, K)
    A = torch.randn(a.gemm, a.gemm, device=dev, dtype=torch.bfloat16)
    B = torch.randn(a.gemm, a.gemm, device=dev, dtype=torch.bfloat16)

    def ep_step():
        buf.dispatch(x, handle=handle, config=cfg)
        buf.combine(cin, handle, config=cfg)

    def gemms(k=4):
        for _ in range(k):
            torch.mm(A, B)

    def timed(ep_stream, gemm_stream, with_ep, with_gemm):
        """mean device time of the EP step and of the GEMM batch when they are enqueued together"""
        t_ep, t_mm = [], []
        for it in range(a.iters + 3):
            torch.cuda.synchronize()
            if n > 1:
                dist.barrier()
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if with_gemm:
                with torch.cuda.stream(gemm_stream):
                    evs[2].record()
                    gemms()
                    evs[3].record()
            if with_ep:
                with torch.cuda.stream(ep_stream):
                    evs[0].record()
                    ep_step()
                    evs[1].record()
            torch.cuda.synchronize()
            if it >= 3:
                if with_ep:
                    t_ep.append(evs[0].elapsed_time(evs[1]))
                if with_gemm:
                    t_mm.append(evs[2].elapsed_time(evs[3]))
        return (mx(sum(t_ep) / len(t_ep)) * 1e3 if t_ep else None, mx(sum(t_mm) / len(t_mm)) * 1e3 if t_mm else None)

    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    out = {}
    out["ep_alone_us"], _ = timed(s1, s2, True, False)
    _, out["gemm_alone_us"] = timed(s1, s2, False, True)
    out["plain_streams_ep_us"], out["plain_streams_gemm_us"] = timed(s1, s2, True, True)
    buf.use_sm_partition(part)  # the buffer's communication stream is now the partition's stream
    ps, rs = buf.get_comm_stream(), rest.stream()
    out["ep_on_partition_alone_us"], _ = timed(ps, rs, True, False)
    _, out["gemm_on_rest_alone_us"] = timed(ps, rs, False, True)
    out["partitioned_ep_us"], out["partitioned_gemm_us"] = timed(ps, rs, True, True)
    out["partition_sms"], out["rest_sms"], out["ep_num_sms"] = part.sm_count, rest.sm_count, cfg.num_sms
    res["ep_vs_gemm"] = out
    buf.use_sm_partition(None)
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            with open(a.out, "w") as f:
                json.dump(res, f, indent=1)
    if n > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
