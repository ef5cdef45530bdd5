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


def _hook_worker(rank, world, port, q):
    import os

    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uccl_b200 import Communicator
    from uccl_b200.parallel.ddp import wrap_ddp

    comm = Communicator.from_torch_dist(None, host=True, heap_bytes=128 << 20, stage_bytes=2 << 20)
    res = []
    for compress in (False, True):
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8))
        ddp = wrap_ddp(model, comm, compress=compress)
        x = torch.full((4, 16), float(rank + 1))
        ddp(x).sum().backward()
        res.append([p.grad.clone() for p in model.parameters()])
    # reference: average of the per-rank local gradients
    torch.manual_seed(0)
    ref_model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.ReLU(), torch.nn.Linear(32, 8))
    grads = []
    for r in range(world):
        ref_model.zero_grad()
        ref_model(torch.full((4, 16), float(r + 1))).sum().backward()
        grads.append([p.grad.clone() for p in ref_model.parameters()])
    exp = [sum(g[i] for g in grads) / world for i in range(len(grads[0]))]
    ok_plain = all(torch.allclose(a, b, rtol=1e-5, atol=1e-6) for a, b in zip(res[0], exp))
    ok_bf16 = all(torch.allclose(a, b, rtol=2e-2, atol=2e-2) for a, b in zip(res[1], exp))
    q.put((rank, ok_plain, ok_bf16))
    dist.destroy_process_group()


def test_ddp_comm_hooks_world2_cpu():
    """DDP over gloo with the gradient reduction swapped for our all-reduce (fused averaging) and for the
    bf16-on-the-wire variant whose widening cast is fused into the reduction."""
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_hook_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=180) for _ in range(2))
    [p.join(30) for p in ps]
    assert got == [(0, True, True), (1, True, True)]


def _resnet_worker(rank, world, path, q):
    import sys

    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    torch.set_num_threads(2)
    import torch.distributed as dist

    import uccl_b200.parallel.pg  # noqa: F401

    # the example calls init_process_group("uccl_b200") itself: give it a file store through the env-free API
    orig = dist.init_process_group

    def init(backend, **kw):
        return orig(backend, rank=rank, world_size=world, store=dist.FileStore(path, world))

    dist.init_process_group = init
    import ddp_train

    loss = ddp_train.main(["--cpu", "--steps", "2", "--warmup", "1", "--batch", "4"])
    q.put((rank, loss))


def test_resnet_ddp_example_on_cpu():
    """examples/ddp_train.py (ResNet-18, synthetic CIFAR shapes) for a few steps over the torch backend on the
    host fabric: the BASELINE 'DDP through the drop-in' flow, runnable without a GPU."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "store")
        ps = [ctx.Process(target=_resnet_worker, args=(r, 2, path, q)) for r in range(2)]
        [p.start() for p in ps]
        got = dict(q.get(timeout=300) for _ in range(2))
        [p.join(60) for p in ps]
    assert all(v == v and v < 20 for v in got.values()), got  # finite, sane cross-entropy
