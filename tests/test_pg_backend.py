"""The "uccl_b200" torch.distributed backend over the host fabric (world_size = 2, CPU)."""
import multiprocessing as mp
import os
import tempfile

import torch


def _worker(rank, world, path, q):
    import torch.distributed as dist

    import uccl_b200.parallel.pg  # noqa: F401  (registers the backend)

    torch.set_num_threads(1)
    store = dist.FileStore(path, world)
    dist.init_process_group("uccl_b200", rank=rank, world_size=world, store=store)
    x = torch.arange(1000, dtype=torch.float32) + rank
    dist.all_reduce(x)
    ok1 = torch.equal(x, sum(torch.arange(1000, dtype=torch.float32) + r for r in range(world)))
    g = torch.empty(world * 4, dtype=torch.int64)
    dist.all_gather_into_tensor(g, torch.full((4,), rank, dtype=torch.int64))
    ok2 = g.view(world, 4)[:, 0].tolist() == list(range(world))
    outs = [torch.empty(3) for _ in range(world)]
    dist.all_gather(outs, torch.full((3,), float(rank)))
    ok3 = [o[0].item() for o in outs] == [float(r) for r in range(world)]
    b = torch.full((5,), float(rank))
    dist.broadcast(b, src=1)
    ok4 = bool((b == 1.0).all())
    rs = torch.empty(10)
    dist.reduce_scatter_tensor(rs, torch.ones(world * 10) * (rank + 1))
    ok5 = bool((rs == sum(range(1, world + 1))).all())
    # DDP through the backend (gradient averaging on the native allreduce path)
    torch.manual_seed(0)
    model = torch.nn.Linear(8, 4)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    inp = torch.full((2, 8), float(rank + 1))
    ddp(inp).sum().backward()
    gsum = model.weight.grad.clone()
    exp = torch.full((4, 8), sum(2.0 * (r + 1) for r in range(world)) / world)
    ok6 = torch.allclose(gsum, exp)
    # point-to-point + all_to_all (equal and unequal splits) + reduce
    peer = 1 - rank
    t = torch.full((7,), float(rank))
    if rank == 0:
        dist.send(t, peer)
        dist.recv(t, peer)
    else:
        r_ = torch.empty(7)
        dist.recv(r_, peer)
        dist.send(t, peer)
        t = r_
    ok7 = bool((t == peer).all()) if rank == 0 else bool((t == 0).all())
    a_out = torch.empty(world * 3)
    dist.all_to_all_single(a_out, torch.arange(world * 3, dtype=torch.float32) + 10 * rank)
    ok8 = a_out.tolist() == [float(10 * s + 3 * rank + i) for s in range(world) for i in range(3)]
    splits_in = [1, 3] if rank == 0 else [2, 1]
    splits_out = [1, 2] if rank == 0 else [3, 1]
    v_in = torch.arange(sum(splits_in), dtype=torch.float32) + 100 * rank
    v_out = torch.empty(sum(splits_out))
    dist.all_to_all_single(v_out, v_in, output_split_sizes=splits_out, input_split_sizes=splits_in)
    exp_v = [0.0, 100.0, 101.0] if rank == 0 else [1.0, 2.0, 3.0, 102.0]
    ok9 = v_out.tolist() == exp_v
    red = torch.full((4,), float(rank + 1))
    dist.reduce(red, dst=1, op=dist.ReduceOp.SUM)
    ok10 = rank != 1 or bool((red == 3.0).all())
    dist.barrier()
    q.put((rank, ok1, ok2, ok3, ok4, ok5, ok6 and ok7 and ok8 and ok9 and ok10))
    dist.destroy_process_group()


def test_pg_backend_world2_cpu():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "store")
        ps = [ctx.Process(target=_worker, args=(r, 2, path, q)) for r in range(2)]
        [p.start() for p in ps]
        got = sorted(q.get(timeout=180) for _ in range(2))
        [p.join(30) for p in ps]
    assert got == [(0, True, True, True, True, True, True), (1, True, True, True, True, True, True)]
