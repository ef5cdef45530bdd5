#!/usr/bin/env python
"""Data-parallel ResNet training on synthetic CIFAR-shaped data, gradients reduced by the native
allreduce (counterpart of the reference's examples/ddp_train.py, which needs the CIFAR-10
download and goes through NCCL + the UCCL net plugin).

  torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/ddp_train.py
  python examples/ddp_train.py --backend nccl         # baseline for comparison
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist
import torch.nn.functional as F


def main(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--backend", default="uccl_b200", choices=["uccl_b200", "nccl", "hook"])
    p.add_argument("--model", default="resnet18", choices=["resnet18", "resnet50"])
    p.add_argument("--batch", type=int, default=128)
    p.add_argument("--steps", type=int, default=30)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--cpu", action="store_true", help="CPU reference backend (host fabric); for smoke-testing the flow")
    p.add_argument("--sym-buckets", action="store_true",
                   help="uccl_b200 / hook: allocate DDP's gradient buckets in the symmetric heap (zero-copy NVLS "
                        "all-reduce instead of the staged kernels) -- the torch-side analogue of ncclMemAlloc")
    p.add_argument("--json", default=None, help="append the result as one JSON line to this file (rank 0)")
    args = p.parse_args(argv)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.cpu:
        dev = torch.device("cpu")
        assert args.backend == "uccl_b200", "--cpu runs the uccl_b200 torch backend over the host fabric"
    else:
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
    sync = (lambda: None) if args.cpu else torch.cuda.synchronize
    if args.backend == "uccl_b200":
        import uccl_b200.parallel.pg  # noqa: F401

        dist.init_process_group("uccl_b200")
    else:
        dist.init_process_group("nccl", device_id=dev)
    from uccl_b200.models import resnet18, resnet50

    small = args.model == "resnet18"
    model = (resnet18(10, True) if small else resnet50(1000, False)).to(dev).to(memory_format=torch.channels_last)
    import contextlib

    comm = None
    if args.backend == "hook":
        from uccl_b200 import Communicator

        comm = Communicator.from_torch_dist(heap_bytes=2 << 30)
    elif args.backend == "uccl_b200" and not args.cpu:
        comm = uccl_b200.parallel.pg.default_communicator()
    pool_ctx = comm.use_mem_pool() if (args.sym_buckets and comm is not None) else contextlib.nullcontext()
    with pool_ctx:  # DDP allocates its flat gradient buckets in the constructor
        ddp = torch.nn.parallel.DistributedDataParallel(
            model, device_ids=[local] if args.backend != "uccl_b200" else None, gradient_as_bucket_view=True)
    if args.backend == "hook":
        from uccl_b200.parallel.ddp import allreduce_hook

        ddp.register_comm_hook(None, allreduce_hook(comm))
    opt = torch.optim.SGD(ddp.parameters(), lr=0.05, momentum=0.9)
    res = 32 if small else 224
    x = torch.randn(args.batch, 3, res, res, device=dev).to(memory_format=torch.channels_last)
    y = torch.randint(0, 10 if small else 1000, (args.batch,), device=dev)
    t0 = None
    for step in range(args.warmup + args.steps):
        if step == args.warmup:
            sync()
            dist.barrier()
            t0 = time.perf_counter()
        with torch.autocast("cpu" if args.cpu else "cuda", dtype=torch.bfloat16, enabled=not args.cpu):
            loss = F.cross_entropy(ddp(x), y)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
    sync()
    dt = time.perf_counter() - t0
    if rank == 0:
        tag = args.backend + ("+sym" if args.sym_buckets else "")
        print(f"[{tag}] {args.model} x{world}: {args.steps * args.batch * world / dt:.0f} img/s, "
              f"{dt / args.steps * 1e3:.2f} ms/step, final loss {loss.item():.4f}")
        if args.json:
            import json

            with open(args.json, "a") as f:
                f.write(json.dumps({"backend": tag, "model": args.model, "gpus": world, "batch_per_gpu": args.batch,
                                    "img_per_s": args.steps * args.batch * world / dt,
                                    "ms_per_step": dt / args.steps * 1e3, "loss": float(loss.item())}) + "\n")
    dist.destroy_process_group()
    return float(loss.item())


if __name__ == "__main__":
    main()
