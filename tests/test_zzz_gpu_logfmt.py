"""GPU test of `low_latency_combine(use_logfmt=True)`: the reference's LogFMT-10 *simulated cast*
(ep/src/internode_ll.cu:934-995) applied by `ep_ll_pack_logfmt_kernel` while the expert outputs are brought into the
symmetric combine buffer, against the fp32 PyTorch definition `uccl_b200.ep.utils.logfmt10_simulate`.
(Sorted last on purpose: the newest kernel of the tree runs after everything else.)"""
import pytest
import torch

from test_gpu_ep import _ll_setup, make_inputs, run_threads
from uccl_b200.ep.utils import logfmt10_simulate

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("n", [2, 4])
@pytest.mark.parametrize("in_place", [False, True])
def test_low_latency_combine_logfmt(n, in_place):
    T, H, K, M = 40, 1024, 4, 64
    E = n * 2
    E_local = E // n
    bufs = _ll_setup(n, M, H)
    xs, idxs, ws = make_inputs(n, T, H, K, E, seed=900 + n)
    xs = [(x.float() * 0.05).to(torch.bfloat16) for x in xs]  # |x| <= 1 almost everywhere: the grid applies

    def fn(b):
        r, dev = b.rank, b.device
        x, idx, w = xs[r].to(dev), idxs[r].to(dev), ws[r].to(dev)
        rx, cnt, handle, _, _ = b.low_latency_dispatch(x, idx, M, E, use_fp8=False)
        torch.cuda.current_stream().synchronize()
        if in_place:
            eo = b.get_next_low_latency_combine_buffer(handle)  # the symmetric buffer itself: cast happens in place
        else:
            eo = torch.zeros(E_local, n * M, H, dtype=torch.bfloat16, device=dev)
        for el in range(E_local):
            c = int(cnt[el])
            eo[el, :c] = rx[el, :c]
        with pytest.raises(ValueError):
            b.low_latency_combine(eo, idx, w, handle, use_logfmt=True, zero_copy=True)
        out_l, _, _ = b.low_latency_combine(eo, idx, w, handle, use_logfmt=True)
        torch.cuda.current_stream().synchronize()
        if in_place:  # refill: the buffer now holds the cast rows
            for el in range(E_local):
                c = int(cnt[el])
                eo[el, :c] = rx[el, :c]
        out, _, _ = b.low_latency_combine(eo, idx, w, handle)
        torch.cuda.current_stream().synchronize()
        return dict(out_l=out_l.float().cpu(), out=out.float().cpu())

    res = run_threads(bufs, fn)
    for r in range(n):
        wsum = torch.where(idxs[r] >= 0, ws[r], torch.zeros_like(ws[r])).sum(1)
        q = logfmt10_simulate(xs[r]).float()
        exp_l = q * wsum[:, None]
        exp = xs[r].float() * wsum[:, None]
        assert torch.allclose(res[r]["out"], exp, rtol=2e-2, atol=1e-3)
        err = (res[r]["out_l"] - exp_l).abs()
        # device log2 / exp2 may land a value that sits on a grid boundary in the neighbouring code: one step (< 4.5 %)
        assert bool((err <= 0.06 * exp_l.abs() + 1e-3).all())
        assert float((err <= 2e-2 * exp_l.abs() + 1e-3).float().mean()) > 0.995
        assert not torch.allclose(res[r]["out_l"], exp, rtol=1e-3, atol=1e-4)  # the grid was applied


def test_raw_buffer_views_and_reset():
    """get_local_buffer_tensor / reset_rdma_buffer / connect_atomic_buffer of the CUDA Buffer (reference:
    ep/bench/buffer.py:213-221,606-647); low-latency traffic still works after a reset."""
    n, T, H, K, M = 2, 16, 1024, 2, 64
    E = n * 2
    bufs = _ll_setup(n, M, H)
    xs, idxs, ws = make_inputs(n, T, H, K, E, seed=77)

    def fn(b):
        dev = b.device
        t = b.get_local_buffer_tensor(torch.bfloat16)
        assert t.is_cuda and t.numel() * 2 == int(b.runtime.arena_area_bytes)
        v = b.get_local_buffer_tensor(torch.float32, torch.Size([4, 8]), offset=t.numel() // 2 - 64)
        v.fill_(1.5)
        again = b.get_local_buffer_tensor(torch.float32, torch.Size([32]), offset=t.numel() // 2 - 64)
        assert bool((again == 1.5).all())
        x, idx, w = xs[b.rank].to(dev), idxs[b.rank].to(dev), ws[b.rank].to(dev)
        rx, cnt, handle, _, _ = b.low_latency_dispatch(x, idx, M, E, use_fp8=False)
        torch.cuda.current_stream().synchronize()
        r = b.get_local_buffer_tensor(torch.uint8, use_rdma_buffer=True)
        assert r.numel() == int(b.runtime.ll_nbytes) and r.numel() > 0
        with pytest.raises(TypeError):
            b.connect_atomic_buffer(None)
        return int(cnt.sum())

    first = run_threads(bufs, fn)

    def after_reset(b):
        b.reset_rdma_buffer()
        torch.cuda.current_stream().synchronize()
        return 0

    run_threads(bufs, after_reset)

    def roundtrip(b):
        dev = b.device
        x, idx, w = xs[b.rank].to(dev), idxs[b.rank].to(dev), ws[b.rank].to(dev)
        rx, cnt, handle, _, _ = b.low_latency_dispatch(x, idx, M, E, use_fp8=False)
        eo = b.get_next_low_latency_combine_buffer(handle)
        for el in range(E // n):
            c = int(cnt[el])
            eo[el, :c] = rx[el, :c]
        out, _, _ = b.low_latency_combine(eo, idx, w, handle)
        torch.cuda.current_stream().synchronize()
        return out.float().cpu(), int(cnt.sum())

    res = run_threads(bufs, roundtrip)
    for r in range(n):
        wsum = torch.where(idxs[r] >= 0, ws[r], torch.zeros_like(ws[r])).sum(1)
        assert torch.allclose(res[r][0], xs[r].float() * wsum[:, None], rtol=2e-2, atol=2e-2)
        assert res[r][1] == first[r]
