"""One rank of tests/test_ukernel.py::test_functional_api_standalone_world: the call sequence of the reference's
experimental/ukernel/py/test_collective.py (stand-alone world through the exchanger, no torch.distributed) on host
memory."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import uccl_b200.compat

uccl_b200.compat.install_ukernel()
import ukernel_ccl as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    pg = dist.init_process_group(backend="ukernel", rank=rank, world_size=world, gpu_id=-1,
                                 exchanger_ip="127.0.0.1", exchanger_port=int(os.environ["EXCHANGER_PORT"]),
                                 transport="auto", device_task_capacity=4096, heap_bytes=256 << 20, stage_bytes=8 << 20)
    assert dist.is_initialized() and dist.get_rank() == rank and dist.get_world_size(pg) == world
    x = torch.arange(0, 1024 * world + 1, dtype=torch.float32) + rank * 1000
    work = dist.all_reduce(x, group=pg, async_op=True, tile_bytes=64 << 10, num_flows=2)
    work.wait()
    exp = torch.arange(0, 1024 * world + 1, dtype=torch.float32) * world + 1000 * sum(range(world))
    assert torch.equal(x, exp), x[:8]
    send = torch.arange(0, 12 * world, dtype=torch.float32) + rank * 10000
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send, group=pg, tile_bytes=64 << 10, num_flows=2)
    for src in range(world):
        assert torch.equal(recv[12 * src:12 * src + 12], torch.arange(12 * rank, 12 * rank + 12, dtype=torch.float32) + src * 10000)
    base = 4
    isp = [base + ((rank + peer) % 2) for peer in range(world)]
    osp = [base + ((src + rank) % 2) for src in range(world)]
    send_v = torch.empty(sum(isp), dtype=torch.float32)
    cur = 0
    for dst, n in enumerate(isp):
        send_v[cur:cur + n] = rank * 10000 + dst * 100
        cur += n
    recv_v = torch.empty(sum(osp), dtype=torch.float32)
    dist.all_to_all_single(recv_v, send_v, output_split_sizes=osp, input_split_sizes=isp, group=pg, tile_bytes=64 << 10,
                           num_flows=2)
    cur = 0
    for src, n in enumerate(osp):
        assert bool((recv_v[cur:cur + n] == src * 10000 + rank * 100).all())
        cur += n
    dist.barrier(group=pg)
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
