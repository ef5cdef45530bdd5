#!/usr/bin/env python
"""Expert-parallel MoE training step on top of `uccl_b200.ep` (the consumer shape Megatron-style trainers
have): router -> differentiable dispatch -> local experts -> differentiable combine, data-parallel
gradient averaging of the router through the communicator, expert weights stay local.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/moe_train.py          # GPUs
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 examples/moe_train.py --cpu    # CPU reference backend
    torchrun --nnodes=2 --nproc-per-node 8 ... examples/moe_train.py --local-size 8                # experts spread over 2 boxes
    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 examples/moe_train.py --cpu --local-size 2   # 2 "boxes" on one machine
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator
from uccl_b200.ep import Buffer
from uccl_b200.models.moe import ExpertParallelMoE


def run(rank: int, world: int, cpu: bool, steps: int = 5, tokens: int = 256, hidden: int = 256, ffn: int = 512,
        experts_per_rank: int = 2, top_k: int = 2, lr: float = 0.5, verbose: bool = True, fixed_batch: bool = False,
        local_size: int = 0):
    """One process of the job; returns the per-step losses (averaged over ranks)."""
    if 0 < local_size < world:
        # the expert-parallel group spans boxes: NVLink (or shared memory) inside a box, datagram rails between;
        # dispatch / combine ride the hierarchical two-hop all-to-all (portable EP path)
        from uccl_b200.parallel import MultiNodeCommunicator

        if not cpu:
            torch.cuda.set_device(rank % torch.cuda.device_count())
        kw = dict(host=True, heap_bytes=256 << 20, stage_bytes=4 << 20) if cpu else dict(heap_bytes=2 << 30, stage_bytes=64 << 20)
        comm = MultiNodeCommunicator.from_torch_dist(local_size, **kw)
        dev = comm.device
    elif cpu:
        comm = Communicator.from_torch_dist(None, host=True, heap_bytes=256 << 20, stage_bytes=4 << 20)
        dev = torch.device("cpu")
    else:
        torch.cuda.set_device(rank % torch.cuda.device_count())
        comm = Communicator.from_torch_dist(None, heap_bytes=2 << 30, stage_bytes=64 << 20)
        dev = comm.device
    buf = Buffer(comm=comm, num_nvl_bytes=512 << 20)
    torch.manual_seed(1234)  # identical router initialisation everywhere (it is data-parallel)
    model = ExpertParallelMoE(hidden, ffn, experts_per_rank * world, top_k, buf).to(dev)
    torch.manual_seed(77 + rank)  # different experts and different data per rank
    with torch.no_grad():
        model.w1.normal_(0, hidden ** -0.5)
        model.w2.normal_(0, ffn ** -0.5)
    teacher = torch.randn(hidden, hidden, device=dev).to(torch.bfloat16) * hidden ** -0.5
    opt = torch.optim.SGD(model.parameters(), lr=lr)
    losses = []
    x_fixed = torch.randn(tokens, hidden, device=dev).to(torch.bfloat16)
    for step in range(steps):
        x = x_fixed if fixed_batch else torch.randn(tokens, hidden, device=dev).to(torch.bfloat16)
        target = (x @ teacher).float()
        y = model(x)
        loss = torch.nn.functional.mse_loss(y.float(), target)
        opt.zero_grad()
        loss.backward()
        # the router is replicated: average its gradient over the ranks (fused 1/N in the all-reduce)
        g = model.router.weight.grad.float().contiguous()
        comm.all_reduce(g, "avg")
        model.router.weight.grad.copy_(g.to(model.router.weight.grad.dtype))
        opt.step()
        l = loss.detach().float().reshape(1).to(dev)
        comm.all_reduce(l, "avg")
        losses.append(float(l.item()))
        if verbose and rank == 0:
            print(f"step {step}: loss {losses[-1]:.4f}")
    return losses


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true", help="CPU reference backend (gloo rendezvous, host communicator)")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--local-size", type=int, default=0, help="ranks per box when the job spans boxes (0: one box)")
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist.init_process_group("gloo" if a.cpu else "cpu:gloo,cuda:nccl")
    losses = run(rank, world, a.cpu, steps=a.steps, local_size=a.local_size)
    if rank == 0:
        print("losses:", [round(v, 4) for v in losses])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
