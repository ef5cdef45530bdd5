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
