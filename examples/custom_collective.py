#!/usr/bin/env python
"""Write a collective of your own, check it on the CPU, run it on the ukernel executor.

The reference's route for custom algorithms is the MSCCL++ DSL -> JSON plan -> `executionKernel` interpreter
(experimental/lite/collective/execution_kernel.hpp:898).  Here: `uccl_b200.ukernel.dsl.Program` -> validate / simulate
(C++) -> JSON -> `Program.run(UkCommunicator, ...)` on the persistent worker (GPU) or its host backend (`--cpu`).

  python examples/custom_collective.py --cpu --ranks 4            # threads, host backend, no GPU needed
  torchrun --nproc-per-node 8 examples/custom_collective.py       # one process per GPU
"""
import argparse
import os
import sys
import threading

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from uccl_b200 import Communicator
from uccl_b200 import ukernel as uk
from uccl_b200.ukernel import dsl
from uccl_b200.ukernel.dsl import In, Out, Program, Scratch


def halving_doubling_pairs(n: int, nbytes: int, elem_size: int) -> Program:
    """A hand-written example: all-reduce on an even number of ranks as (1) pairwise exchange-and-add between
    neighbours 2i / 2i+1, (2) recursive doubling among the even ranks only, (3) even ranks hand the result to their odd
    neighbour.  Halves the number of ranks in the log-step phase -- the shape of a two-level (e.g. per-die) scheme."""
    assert n % 2 == 0 and (n // 2) & (n // 2 - 1) == 0
    rounds = (n // 2).bit_length() - 1
    p = Program(f"pair_reduce_doubling_{n}", n, 1, nbytes, nbytes, (rounds + 1) * nbytes, elem_size)
    for r in range(n):
        p.copy(r, Out(0), In(0), nbytes)
    for e in range(0, n, 2):  # (1) the odd neighbour contributes
        p.send(e + 1, e, Scratch(0), Out(0), nbytes)
        p.reduce(e, Out(0), Out(0), Scratch(0), nbytes)
    for k in range(rounds):   # (2) recursive doubling over the even ranks
        slot = (k + 1) * nbytes
        hs = [p.isend(e, ((e // 2) ^ (1 << k)) * 2, Scratch(slot), Out(0), nbytes) for e in range(0, n, 2)]
        for h in hs:
            p.wait(h)
        for e in range(0, n, 2):
            p.reduce(e, Out(0), Out(0), Scratch(slot), nbytes)
    for e in range(0, n, 2):  # (3) result back to the odd neighbour
        p.send(e, e + 1, Out(0), Out(0), nbytes)
    return p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--ranks", type=int, default=4)
    ap.add_argument("--numel", type=int, default=1 << 16)
    args = ap.parse_args()
    n = args.ranks if args.cpu else int(os.environ.get("WORLD_SIZE", "1"))
    nbytes = args.numel * 4
    programs = [dsl.recursive_doubling_allreduce(n, nbytes, 4, nlanes=2), halving_doubling_pairs(n, nbytes, 4)]
    for p in programs:  # 1. structure, 2. numerics on host memory with the greedy simulator
        p.validate()
        ins = [torch.full((args.numel,), float(r + 1)) for r in range(n)]
        outs = [torch.zeros(args.numel) for _ in range(n)]
        p.simulate(ins, outs, "sum")
        assert all(bool((o == n * (n + 1) / 2).all()) for o in outs), p.name
        print(f"{p.name}: {p.num_ops()} ops, validated + simulated, JSON {len(p.to_json())} bytes")

    def rank_main(comm):
        u = uk.UkCommunicator(comm, nlanes=2, staging_bytes=max(1 << 20, nbytes))
        for p in programs:
            x = torch.full((args.numel,), float(comm.rank + 1), device=comm.device)
            Program.from_json(p.to_json()).run(u, x, op="sum").wait()
            if not comm.is_host:
                torch.cuda.synchronize(comm.device)
            assert bool((x == n * (n + 1) / 2).all()), (p.name, comm.rank)
        u.stop()
        if comm.rank == 0:
            print(f"executed on {'the host backend' if comm.is_host else 'the device worker'}: ok")

    if args.cpu:
        comms = Communicator.local_world(n, host=True, heap_bytes=max(128 << 20, 16 * nbytes), stage_bytes=1 << 20)
        ts = [threading.Thread(target=rank_main, args=(c,)) for c in comms]
        [t.start() for t in ts]
        [t.join() for t in ts]
    else:
        import torch.distributed as dist

        dist.init_process_group("gloo")
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        rank_main(Communicator.from_torch_dist(heap_bytes=max(256 << 20, 16 * nbytes)))
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
