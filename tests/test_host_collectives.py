"""CPU tests of the control plane + host reference collectives (BASELINE config #1:
'NCCL-API allreduce correctness world_size=2 on CPU/gloo (plumbing, no GPU)')."""
import multiprocessing as mp

import pytest
import torch

from helpers import run_host_ranks
from uccl_b200 import Communicator


@pytest.fixture(scope="module")
def world4():
    return Communicator.local_world(4, host=True, heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=20000)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16, torch.int32, torch.int64, torch.float64])
@pytest.mark.parametrize("op", ["sum", "max", "min", "avg"])
def test_host_allreduce(world4, dtype, op):
    n = 4
    count = 70001  # spans several stage chunks for 8-byte types
    gen = torch.Generator().manual_seed(0)
    ins = [(torch.randint(-8, 8, (count,), generator=gen)).to(dtype) for _ in range(n)]
    ref = torch.stack([x.double() for x in ins])
    if op == "sum":
        exp = ref.sum(0)
    elif op == "avg":
        exp = ref.sum(0) / n
        if not dtype.is_floating_point:
            exp = torch.div(ref.sum(0), n, rounding_mode="trunc")
    elif op == "max":
        exp = ref.max(0).values
    else:
        exp = ref.min(0).values

    def fn(c):
        x = ins[c.rank].clone()
        c.all_reduce(x, op)
        return x

    outs = run_host_ranks(world4, fn)
    for o in outs:
        assert torch.allclose(o.double(), exp.to(dtype).double(), rtol=1e-2, atol=1e-2)
        assert torch.equal(o, outs[0])


def test_host_other_collectives(world4):
    n = 4

    def fn(c):
        r = c.rank
        g = torch.empty(n * 33, dtype=torch.float32)
        c.all_gather(g, torch.full((33,), float(r)))
        rs = torch.empty(50, dtype=torch.float32)
        c.reduce_scatter(rs, torch.arange(n * 50, dtype=torch.float32) * (r + 1), "sum")
        b = torch.full((17,), float(r))
        c.broadcast(b, root=2)
        red = torch.full((9,), float(r + 1))
        c.reduce(red, root=1, op="prod")
        a2a = torch.empty(n * 3, dtype=torch.int32)
        c.all_to_all(a2a, (torch.arange(n * 3, dtype=torch.int32) + 100 * r))
        c.barrier()
        return g, rs, b, red, a2a

    outs = run_host_ranks(world4, fn)
    for r, (g, rs, b, red, a2a) in enumerate(outs):
        assert torch.equal(g.view(n, 33)[:, 0], torch.arange(n, dtype=torch.float32))
        exp_rs = torch.arange(n * 50, dtype=torch.float32).view(n, 50)[r] * sum(range(1, n + 1))
        assert torch.equal(rs, exp_rs)
        assert torch.equal(b, torch.full((17,), 2.0))
        if r == 1:
            assert torch.equal(red, torch.full((9,), 24.0))
        exp_a2a = torch.cat([torch.arange(3, dtype=torch.int32) + 3 * r + 100 * s for s in range(n)])
        assert torch.equal(a2a, exp_a2a)


def test_host_allreduce_fused_scale_and_cast(world4):
    """fp32 gradients reduced with a fused scale and rounded once to bf16 / fp16 (and the reverse widening):
    the host backend follows the CUDA epilogue's arithmetic -- fp32 accumulate, one rounding."""
    n = len(world4)
    g = torch.Generator().manual_seed(1)
    ins = [torch.randn(3001, generator=g) for _ in range(n)]
    ref = torch.stack(ins).sum(0) * 0.5

    def fn(c):
        outs = {}
        for dt in (torch.bfloat16, torch.float16):
            o = torch.zeros(3001, dtype=dt)
            c.all_reduce(ins[c.rank].clone(), "sum", out=o, scale=0.5)
            outs[dt] = o
        w = torch.zeros(3001)
        c.all_reduce(ins[c.rank].to(torch.bfloat16), "avg", out=w)  # bf16 in, fp32 out
        return outs, w

    res = run_host_ranks(world4, fn)
    wide_ref = torch.stack([x.to(torch.bfloat16).float() for x in ins]).sum(0) / n
    for outs, w in res:
        assert torch.equal(outs[torch.bfloat16], ref.to(torch.bfloat16))
        assert torch.equal(outs[torch.float16], ref.to(torch.float16))
        assert torch.allclose(w, wide_ref, rtol=1e-6, atol=1e-6)


def test_host_alltoallv_multi_round(world4):
    """Variable-size all-to-all incl. empty pairs and messages larger than the per-destination stage
    slice (several rounds, ranks finishing in different rounds)."""
    n = len(world4)
    # rank s sends cnt[s][d] int64 values to rank d; one pair is far larger than the others
    cnt = [[(s * 7 + d * 3) % 5 * 1000 for d in range(n)] for s in range(n)]
    cnt[1][2] = 300000  # 2.4 MB through a 256 KiB slice (1 MiB stage / 4 ranks)
    cnt[3][0] = 0

    def fn(c):
        r = c.rank
        x = torch.cat([torch.full((cnt[r][d],), 1000 * r + d, dtype=torch.int64) for d in range(n)] + [torch.empty(0, dtype=torch.int64)])
        out = torch.zeros(sum(cnt[s][r] for s in range(n)), dtype=torch.int64)
        c.all_to_all_v(out, x, cnt[r], [cnt[s][r] for s in range(n)])
        return out

    outs = run_host_ranks(world4, fn)
    for r, o in enumerate(outs):
        exp = torch.cat([torch.full((cnt[s][r],), 1000 * s + r, dtype=torch.int64) for s in range(n)])
        assert torch.equal(o, exp)


def test_host_grouped_send_recv(world4):
    """Grouped point-to-point on the host backend: a ring step with a message larger than a mailbox chunk,
    then an all-to-all pattern incl. self, twice (sequence numbers persist across groups)."""
    n = len(world4)

    def fn(c):
        r = c.rank
        big = torch.arange(3_000_000, dtype=torch.float32) + r          # 12 MB > the 4 MiB mailbox
        got = torch.zeros(3_000_000)
        c.batch_send_recv([("recv", got, (r - 1) % n), ("send", big, (r + 1) % n)])
        outs = []
        for rep in range(2):
            src = torch.stack([torch.full((1000,), float(100 * r + p + rep)) for p in range(n)])
            dst = torch.zeros(n, 1000)
            ops = []
            for p in range(n):
                ops += [("send", src[p], p), ("recv", dst[p], p)]
            c.batch_send_recv(ops)
            outs.append(dst)
        return got, outs

    for r, (got, outs) in enumerate(run_host_ranks(world4, fn)):
        assert torch.equal(got, torch.arange(3_000_000, dtype=torch.float32) + (r - 1) % n)
        for rep, dst in enumerate(outs):
            for p in range(n):
                assert bool((dst[p] == 100 * p + r + rep).all())


def test_symmetric_heap_alloc(world4):
    c = world4[0]
    free0 = c.native.heap_free_bytes
    a = c.empty(1000, dtype=torch.float32)
    b = c.empty(3, 5, dtype=torch.bfloat16)
    assert c.is_symmetric(a) and c.is_symmetric(b)
    assert not c.is_symmetric(torch.empty(4))
    assert a.data_ptr() % 256 == 0 and b.data_ptr() % 256 == 0
    a.fill_(3.0)
    assert a.sum().item() == 3000.0
    del a, b
    import gc

    gc.collect()
    assert c.native.heap_free_bytes == free0


def _mp_worker(rank, n, uid, q):
    torch.set_num_threads(1)
    c = Communicator.init(uid, rank, n, host=True, heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=20000)
    x = torch.arange(5000, dtype=torch.float32) + rank
    c.all_reduce(x, "sum")
    exp = sum(torch.arange(5000, dtype=torch.float32) + r for r in range(n))
    q.put((rank, bool(torch.equal(x, exp))))


def test_multiprocess_bootstrap_world2():
    """Two real processes: TCP rendezvous + shm symmetric heaps + allreduce (world_size=2)."""
    ctx = mp.get_context("spawn")
    uid = Communicator.create_unique_id()
    q = ctx.Queue()
    ps = [ctx.Process(target=_mp_worker, args=(r, 2, uid, q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in ps]
    assert got == [(0, True), (1, True)]


def test_lost_peer_is_a_diagnosable_timeout():
    """Failure detection (reference: spin timeouts that name the site instead of hanging): a rank
    whose peer never shows up gets an error that names the missing rank -- not a hang."""
    import time

    comms = Communicator.local_world(2, host=True, heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=300)
    t0 = time.time()
    with pytest.raises(RuntimeError) as ei:
        comms[0].all_reduce(torch.ones(8), "sum")  # rank 1 never calls
    assert 0.2 < time.time() - t0 < 5.0
    assert "timeout" in str(ei.value) and "rank 1" in str(ei.value)
