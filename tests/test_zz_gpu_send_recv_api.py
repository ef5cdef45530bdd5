"""Python-level grouped send/recv on the GPU (`Communicator.batch_send_recv`): same kernel and the same
patterns as the C++ NCCL-API test (ring step with a multi-chunk message, all-to-all incl. self), through
the pybind path.  Collected last on purpose (newest, least exercised binding)."""
import pytest
import torch

from helpers import get_world, run_ranks

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.mark.parametrize("n", [2, 4])
def test_batch_send_recv_ring_and_alltoall(n):
    comms = get_world(n)
    M = (3 << 20) // 4 + 16   # > one 512 KiB staging slot per block
    C = 40000

    def prepare(c):
        r = c.rank
        s_buf = (torch.arange(M, dtype=torch.float32) % 251 + 1000 * r).to(c.device)
        r_buf = torch.zeros(M, device=c.device)
        a_in = torch.stack([torch.full((C,), float(100 * r + p)) for p in range(n)]).to(c.device)
        a_out = torch.zeros(n, C, device=c.device)
        return s_buf, r_buf, a_in, a_out

    def launch(c, st):
        s_buf, r_buf, a_in, a_out = st
        r = c.rank
        c.batch_send_recv([("send", s_buf, (r + 1) % n), ("recv", r_buf, (r - 1) % n)])
        ops = []
        for p in range(n):
            ops += [("send", a_in[p], p), ("recv", a_out[p], p)]
        c.batch_send_recv(ops)

    outs = run_ranks(comms, prepare, launch)
    for r, (_, r_buf, _, a_out) in enumerate(outs):
        src = (r - 1) % n
        assert torch.equal(r_buf.cpu(), torch.arange(M, dtype=torch.float32) % 251 + 1000 * src)
        for p in range(n):
            assert bool((a_out[p] == 100 * p + r).all())


def test_moe_layer_forward_backward_on_gpu():
    """EP=1 training step of models.ExpertParallelMoE through the CUDA dispatch/combine kernels
    (ep.autograd) against a dense PyTorch evaluation of the same experts."""
    import torch.nn.functional as F

    from uccl_b200 import Communicator
    from uccl_b200.ep import Buffer
    from uccl_b200.models.moe import ExpertParallelMoE

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    T, H, FFN, E, K = 64, 256, 128, 4, 2
    comm = Communicator.local_world(1, devices=[0], heap_bytes=192 << 20, stage_bytes=8 << 20, timeout_ms=5000)[0]
    torch.manual_seed(3)
    m = ExpertParallelMoE(H, FFN, E, K, Buffer(comm=comm, num_nvl_bytes=64 << 20)).to(dev)
    x = torch.randn(T, H, device=dev).to(torch.bfloat16).requires_grad_(True)
    g = torch.randn(T, H, device=dev).to(torch.bfloat16)
    y = m(x)
    (y.float() * g.float()).sum().backward()
    torch.cuda.synchronize()
    # dense reference in fp32 with the same parameters
    w1, w2, rw = (p.detach().float().clone().requires_grad_(True) for p in (m.w1, m.w2, m.router.weight))
    xr = x.detach().float().clone().requires_grad_(True)
    w, idx = torch.topk(F.softmax(xr @ rw.T, dim=-1), K, dim=-1)
    yr = torch.zeros(T, H, device=dev)
    for e in range(E):
        sel = idx == e
        rows = sel.any(1).nonzero().flatten()
        if rows.numel():
            yr = yr.index_add(0, rows, (F.silu(xr[rows] @ w1[e]) @ w2[e]) * (w * sel).sum(1)[rows][:, None])
    (yr * g.float()).sum().backward()

    def close(a, b, tol):
        a, b = a.float(), b.float()
        return (a - b).abs().max().item() <= tol * max(1.0, b.abs().max().item())

    assert close(y, yr.detach(), 5e-2)
    assert close(x.grad, xr.grad, 8e-2)
    assert close(m.w1.grad, w1.grad, 8e-2) and close(m.w2.grad, w2.grad, 8e-2)
    assert close(m.router.weight.grad, rw.grad, 1e-1)
