"""torchrun worker for the multi-process GPU tests (one process per GPU)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    what = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if what == "pg":
        import uccl_b200.parallel.pg  # noqa: F401

        dist.init_process_group("uccl_b200")
        x = torch.full((1 << 20,), float(rank + 1), device=dev)
        dist.all_reduce(x)
        assert x[0].item() == world * (world + 1) / 2
        g = torch.empty(world * 8, device=dev)
        dist.all_gather_into_tensor(g, torch.full((8,), float(rank), device=dev))
        assert g.view(world, 8)[:, 0].tolist() == [float(r) for r in range(world)]
        rs = torch.empty(1000, device=dev)
        dist.reduce_scatter_tensor(rs, torch.ones(world * 1000, device=dev) * (rank + 1))
        assert rs[0].item() == world * (world + 1) / 2
        b = torch.full((33,), float(rank), device=dev)
        dist.broadcast(b, src=world - 1)
        assert b[0].item() == world - 1
        # DDP over the backend
        torch.manual_seed(0)
        model = torch.nn.Linear(64, 32).to(dev)
        ddp = torch.nn.parallel.DistributedDataParallel(model)
        ddp(torch.full((4, 64), float(rank + 1), device=dev)).sum().backward()
        exp = 4.0 * sum(r + 1 for r in range(world)) / world
        assert abs(model.weight.grad[0, 0].item() - exp) < 1e-4, (model.weight.grad[0, 0].item(), exp)
        dist.barrier()
    elif what == "collective":
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
        from uccl_b200 import collective

        collective.init_collective(4, local, heap_bytes=512 << 20)
        nxt, prv = (rank + 1) % world, (rank - 1) % world
        s = torch.full((1 << 18,), float(rank), device=dev)
        r = torch.empty_like(s)
        hs = collective.batch_isend_irecv([collective.P2POp(collective.isend, s, nxt),
                                           collective.P2POp(collective.irecv, r, prv)])
        collective.wait_all(hs)
        torch.cuda.synchronize()
        assert r[0].item() == float(prv) and r[-1].item() == float(prv)
        out = torch.empty(world * 100, device=dev)
        collective.allgather(torch.full((100,), float(rank), device=dev), out)
        torch.cuda.synchronize()
        assert out.view(world, 100)[:, 0].tolist() == [float(q) for q in range(world)]
        t = torch.full((4096,), 1.0, device=dev)
        collective.all_reduce(t, "sum")
        torch.cuda.synchronize()
        assert t[0].item() == world
        collective.finalize_collective()
    elif what == "ep":
        dist.init_process_group("cpu:gloo,cuda:nccl", device_id=dev)
        from uccl_b200.ep import Buffer

        T, H, K, E = 512, 2048, 4, world * 4
        buf = Buffer(dist.group.WORLD, num_nvl_bytes=256 << 20)
        g = torch.Generator(device="cpu").manual_seed(rank)
        x = torch.full((T, H), float(rank + 1), dtype=torch.bfloat16, device=dev)
        idx = torch.rand(T, E, generator=g).topk(K, dim=1).indices.to(torch.int64).to(dev)
        w = torch.ones(T, K, device=dev)
        tpr, _, tpe, in_rank, _ = buf.get_dispatch_layout(idx, E)
        rx, ridx, rw, pe, handle, _ = buf.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank,
                                                   num_tokens_per_expert=tpe, topk_idx=idx, topk_weights=w)
        torch.cuda.synchronize()
        # every received row is constant = source rank + 1 (reference oracle: test_intranode.py:115-118)
        src_rank_of_row = rx[:, 0].float()
        assert bool((rx.float() == src_rank_of_row[:, None]).all())
        cb = buf.get_combine_buffer(rx.size(0), H, K)
        cb.copy_(rx)
        out, _, _ = buf.combine(cb, handle)
        torch.cuda.synchronize()
        fan = in_rank.sum(1).float()
        assert torch.allclose(out.float(), (rank + 1) * fan[:, None].expand(T, H), rtol=1e-2, atol=1e-2)
    dist.barrier()
    if rank == 0:
        print(f"mp_worker {what}: OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
