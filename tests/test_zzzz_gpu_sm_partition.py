"""GPU test of SM partitions (CUDA green contexts; csrc/common/sm_partition.cc, reference probe:
experimental/misc/cuda_greenctx.cu): the split is physical -- kernels on a partition's stream only land on its SMs,
the two sides are disjoint -- and torch work on a partition stream computes the same values.  Sorted last."""
import pytest
import torch

from uccl_b200.utils import SmPartition, sm_ids

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_split_is_physical_and_usable():
    ok, why = SmPartition.supported()
    if not ok:
        pytest.skip(why)
    total = torch.cuda.get_device_properties(0).multi_processor_count
    everywhere = sm_ids(blocks=8 * total)
    assert everywhere.numel() > total // 2  # the probe really spreads
    part, rest = SmPartition.split(24)
    assert 24 <= part.sm_count <= 40 and part.total_sms > part.sm_count
    assert rest is not None and rest.sm_count > 0 and part.sm_count + rest.sm_count <= part.total_sms
    a = part.sm_ids()
    b = rest.sm_ids()
    assert 0 < a.numel() <= part.sm_count and 0 < b.numel() <= rest.sm_count
    assert set(a.tolist()).isdisjoint(set(b.tolist()))
    # torch work on the partition
    x = torch.randn(1 << 20, device="cuda")
    y = torch.randn(1 << 20, device="cuda")
    ref = torch.sin(x) * y + x.abs().sum()
    with part:
        z = torch.sin(x) * y + x.abs().sum()
    with rest:
        z2 = torch.sin(x) * y + x.abs().sum()
    torch.cuda.synchronize()
    assert torch.allclose(z, ref, rtol=1e-4, atol=1e-2) and torch.allclose(z2, ref, rtol=1e-4, atol=1e-2)
    assert "SmPartition" in repr(part)
    # tensors that were allocated while a partition stream was current go first, then the partitions
    del z, z2
    torch.cuda.synchronize()
    del part, rest


def test_ep_buffer_on_a_partition():
    """One rank, real kernels: dispatch + combine on a 24-SM partition give the same rows as on an ordinary stream."""
    ok, why = SmPartition.supported()
    if not ok:
        pytest.skip(why)
    from uccl_b200 import Communicator
    from uccl_b200.ep import Buffer, Config

    T, H, K, E = 256, 1024, 4, 8
    comm = Communicator.local_world(1, devices=[0], heap_bytes=512 << 20, timeout_ms=20000)[0]
    buf = Buffer(comm=comm, num_nvl_bytes=64 << 20)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16).cuda()
    idx = torch.rand(T, E, generator=g).topk(K, dim=1).indices.cuda()
    w = torch.rand(T, K, generator=g).cuda()

    def roundtrip():
        tpr, _, tpe, inr, _ = buf.get_dispatch_layout(idx, E)
        rx, ri, rw, pe, h, _ = buf.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=inr, num_tokens_per_expert=tpe,
                                            topk_idx=idx, topk_weights=w, config=Config(24))
        torch.cuda.current_stream().synchronize()
        keep = rx.float().cpu()
        cin = buf.get_combine_buffer(rx.size(0), H, K)
        cin.copy_(rx)
        out, _, _ = buf.combine(cin, h, topk_weights=rw, config=Config(24))
        torch.cuda.synchronize()
        assert torch.equal(out, x)  # EP = 1: dispatch + combine is the identity
        return keep, out.float().cpu()

    base = roundtrip()
    part, rest = SmPartition.split(24)
    buf.use_sm_partition(part)
    assert buf._sms(Config(64)) <= part.sm_count and buf.get_comm_stream().cuda_stream == part.stream(-1).cuda_stream
    with rest:
        got = roundtrip()
    buf.use_sm_partition(None)
    again = roundtrip()
    for a, b in zip(base, got):
        assert torch.equal(a, b)
    for a, b in zip(base, again):
        assert torch.equal(a, b)
    torch.cuda.synchronize()
    del part, rest
