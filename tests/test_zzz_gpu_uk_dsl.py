"""GPU test of user-authored ukernel programs (uccl_b200.ukernel.dsl) on the persistent device worker: the programs
the CPU tests validate / simulate / run on the host backend (tests/test_ukernel.py), executed by `UkComm::run_custom`
on virtual ranks, staged and in place.  (Sorted last: written after the round's GPU budget was spent.)"""
import pytest
import torch

from helpers import get_world
from uccl_b200 import ukernel as uk
from uccl_b200.ukernel import dsl

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


@pytest.mark.parametrize("n", [2, 4])
def test_gpu_dsl_programs(n):
    comms = get_world(n)
    uks = [uk.UkCommunicator(c, nlanes=2, tile_bytes=64 << 10, staging_bytes=1 << 20) for c in comms]
    numel = 50000
    rd = dsl.recursive_doubling_allreduce(n, numel * 4, elem_size=4, nlanes=2)
    bc = dsl.binomial_broadcast(n, numel * 4, root=n - 1, nlanes=2)
    rd.validate(), bc.validate()
    try:
        ins = [torch.randn(numel, generator=torch.Generator().manual_seed(40 + r)) for r in range(n)]
        ref = torch.stack(ins).sum(0)
        state = []
        for c in comms:
            with torch.cuda.device(c.device):
                s = torch.cuda.Stream(device=c.device)
                x = ins[c.rank].to(c.device)
                y = torch.zeros(numel, device=c.device)
                b = torch.full((numel,), float(c.rank), device=c.device)
                warm = x * 1.0  # load the elementwise kernels before any worker spins (lazy module loading)
                state.append((s, x, y, b))
            torch.cuda.synchronize(c.device)
        works = []
        for c, u, (s, x, y, b) in zip(comms, uks, state):  # one stream per virtual rank (see test_gpu_ukcomm_collectives)
            with torch.cuda.device(c.device), torch.cuda.stream(s):
                works.append(rd.run(u, x, y, "sum"))   # out of place, staged
                works.append(bc.run(u, b))             # in place, staged
        for w in works:
            w.wait()
        for c, (s, x, y, b) in zip(comms, state):
            s.synchronize()
            assert torch.equal(x.cpu(), ins[c.rank])
            assert torch.allclose(y.cpu(), ref, rtol=1e-5, atol=1e-4)
            assert bool((b.cpu() == float(n - 1)).all())
        assert uks[0].stats()["ops"] == 2
    finally:
        for u in uks:
            u.stop()
