#!/usr/bin/env python
"""Several communicators alive at once (the reference's examples/multi_pg_test.py stresses
multiple NCCL process groups): world group + split halves, interleaved collectives.

  torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 examples/multi_pg_test.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from uccl_b200 import Communicator


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    world_comm = Communicator.from_torch_dist(heap_bytes=512 << 20)
    half = world // 2
    groups = [dist.new_group(list(range(0, half))), dist.new_group(list(range(half, world)))]
    mine = groups[0] if rank < half else groups[1]
    sub = Communicator.from_torch_dist(mine, heap_bytes=512 << 20) if half >= 1 else None
    for it in range(20):
        a = torch.full((1 << 16,), float(rank), device="cuda")
        b = torch.full((1 << 12,), 1.0, device="cuda")
        world_comm.all_reduce(a, "sum")
        if sub is not None:
            sub.all_reduce(b, "sum")
        torch.cuda.synchronize()
        assert a[0].item() == sum(range(world)), a[0].item()
        assert sub is None or b[0].item() == sub.world_size
    if rank == 0:
        print("multi_pg_test: OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
