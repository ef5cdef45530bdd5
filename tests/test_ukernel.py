"""ukernel: planner (structure + simulated execution), persistent-worker FIFOs and the ukernel
communicator.  Mirrors the reference's test layering (planner / lowering / ring-allreduce simulator /
executor with mock backends: experimental/ukernel/src/ccl/test/unit/test_components.cc:240-635, then
multi-process collectives) -- the CPU half runs the same FIFO protocol with host threads."""
import threading

import pytest
import torch

from uccl_b200 import Communicator
from uccl_b200 import ukernel as uk
from helpers import get_world


# ------------------------------------------------------------------ planner (CPU)
@pytest.mark.parametrize("algo", ["ring", "fullmesh"])
@pytest.mark.parametrize("n", [1, 2, 3, 4, 8])
def test_plan_allreduce_structure(algo, n):
    for nbytes in (16, 48, 4000, (1 << 16) + 32):
        for lanes in (1, 3):
            assert uk.validate("allreduce", nbytes, n, lanes, 4096, 4, algo) == ""
    text, ops = uk.plan("allreduce", 1 << 16, n, 0, nlanes=2, tile_bytes=4096, elem_size=4, algo=algo)
    assert ("ring" if algo == "ring" else "fullmesh") in text
    if n > 1:
        sends = [o for o in ops if o["kind"] == "send"]
        recvs = [o for o in ops if o["kind"] == "recv"]
        assert len(sends) == len(recvs) > 0
        if algo == "ring":  # a ring only ever talks to its two neighbours
            assert {o["peer"] for o in sends} == {1 % n} and {o["peer"] for o in recvs} == {(n - 1) % n}
        else:
            assert {o["peer"] for o in sends} == set(range(1, n))
    for i, o in enumerate(ops):  # dependencies point backwards and stay on the op's lane
        for d in o["deps"]:
            assert d < i and ops[d]["lane"] == o["lane"]


def test_plan_other_collectives_structure():
    for n in (2, 4, 8):
        assert uk.validate("alltoall", 5000, n, 2, 1024) == ""
        assert uk.validate("allgather", 5008, n, 3, 1024) == ""
        assert uk.validate("barrier", 0, n) == ""
    assert uk._uk().select_algo(uk.COLLS["allreduce"], 8, 1 << 20) == uk.ALGOS["fullmesh"]
    # ring needs (n-1) scratch slots per lane, the full mesh n
    assert uk._uk().scratch_bytes(uk.ALGOS["ring"], 8, 2, 4096) == 2 * 7 * 4096
    assert uk._uk().scratch_bytes(uk.ALGOS["fullmesh"], 8, 2, 4096) == 2 * 8 * 4096


@pytest.mark.parametrize("algo", ["ring", "fullmesh"])
@pytest.mark.parametrize("n", [2, 3, 8])
@pytest.mark.parametrize("dtype,op", [(torch.float32, "sum"), (torch.int32, "max"), (torch.bfloat16, "sum")])
def test_simulated_allreduce(algo, n, dtype, op):
    """All ranks' plans executed over host memory by the greedy reference scheduler."""
    for count in (8, 1000, 70000):
        g = torch.Generator().manual_seed(count)
        if dtype.is_floating_point:
            ins = [torch.randn(count, generator=g).to(dtype) for _ in range(n)]
        else:
            ins = [torch.randint(-99, 99, (count,), generator=g).to(dtype) for _ in range(n)]
        outs = [torch.zeros(count, dtype=dtype) for _ in range(n)]
        assert uk.simulate("allreduce", ins, outs, op, nlanes=3, tile_bytes=4096, algo=algo) == ""
        ref = torch.stack([x.double() for x in ins])
        ref = ref.sum(0) if op == "sum" else ref.max(0).values
        tol = dict(rtol=3e-2, atol=0.3) if dtype == torch.bfloat16 else dict(rtol=1e-5, atol=1e-4)
        for o in outs:
            assert torch.allclose(o.double(), ref, **tol)
        for o in outs[1:]:  # every rank ends with the very same bits
            assert torch.equal(o, outs[0])
        # in place
        bufs = [x.clone() for x in ins]
        assert uk.simulate("allreduce", bufs, bufs, op, nlanes=2, tile_bytes=2048, algo=algo) == ""
        for b in bufs:
            assert torch.equal(b, outs[0]) or torch.allclose(b.double(), ref, **tol)


def test_simulated_alltoall_allgather():
    n, per = 4, 3000
    ins = [torch.arange(n * per, dtype=torch.float32) + 1e5 * r for r in range(n)]
    outs = [torch.zeros(n * per) for _ in range(n)]
    assert uk.simulate("alltoall", ins, outs, nlanes=2, tile_bytes=4096) == ""
    for r in range(n):
        assert torch.equal(outs[r], torch.cat([ins[s][r * per:(r + 1) * per] for s in range(n)]))
    gin = [torch.full((per,), float(r)) for r in range(n)]
    gout = [torch.zeros(n * per) for _ in range(n)]
    assert uk.simulate("allgather", gin, gout, nlanes=3, tile_bytes=1024) == ""
    for o in gout:
        assert torch.equal(o, torch.cat(gin))


# ------------------------------------------------------------------ worker + communicator, host threads
def test_host_worker_tasks():
    w = uk.Worker(device=-1, nlanes=2)
    a = torch.arange(1000, dtype=torch.float32)
    b = torch.ones(1000)
    c = torch.zeros(1000)
    flag = torch.zeros(1, dtype=torch.int64)
    w.copy(0, c.data_ptr(), a.data_ptr(), 4000)
    w.reduce(0, c.data_ptr(), c.data_ptr(), b.data_ptr(), 4000, torch.float32, "sum")
    w.signal(0, flag.data_ptr(), 5)
    t = w.wait_value(1, flag.data_ptr(), 5)  # lane 1 is released by lane 0's signal
    w.wait(1, t)
    w.wait_all()
    assert torch.equal(c, a + 1) and flag.item() == 5
    # more tasks than ring entries: the producer blocks on the consumer, nothing is lost
    acc = torch.zeros(4, dtype=torch.int64)
    one = torch.ones(4, dtype=torch.int64)
    n_tasks = uk._uk().RING_ENTRIES * 2 + 7
    for _ in range(n_tasks):
        w.reduce(0, acc.data_ptr(), acc.data_ptr(), one.data_ptr(), 32, torch.int64, "sum")
    w.wait_all()
    assert acc.tolist() == [n_tasks] * 4
    assert w.stats()["reduces"] == n_tasks + 1 and w.error == 0
    w.stop()


def _run_threads(comms, fn):
    res, errs = [None] * len(comms), []

    def body(c):
        try:
            res[c.rank] = fn(c)
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=body, args=(c,)) for c in comms]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    return res


@pytest.mark.parametrize("n", [2, 4])
def test_host_ukcomm_collectives(n):
    comms = Communicator.local_world(n, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20)

    def fn(c):
        if c.rank == 1:
            c._c.alloc(12345, 256)  # ranks need not allocate symmetrically: region offsets are exchanged
        pg = uk.ProcessGroup(c, nlanes=2, tile_bytes=4096, staging_bytes=64 << 10)
        x = torch.full((50000,), float(c.rank + 1))  # > staging: several segments
        pg.all_reduce(x, "sum")
        y = torch.arange(n * 1000, dtype=torch.float32) + 10000 * c.rank
        z = torch.zeros(n * 1000)
        pg.all_to_all_single(z, y)
        g = torch.zeros(n * 777)
        pg.all_gather_into_tensor(g, torch.full((777,), float(c.rank)))
        pg.barrier()
        xr = torch.full((5000,), float(c.rank + 1))
        w = pg._uk.all_reduce(xr, "max", algo="ring")
        w.wait()
        assert w.is_completed()
        xa = torch.full((64,), float(c.rank))
        pg.all_reduce(xa, "avg")
        st = pg._uk.stats()
        pg.shutdown()
        return x, z, g, xr, xa, st

    for r, (x, z, g, xr, xa, st) in enumerate(_run_threads(comms, fn)):
        assert torch.all(x == n * (n + 1) / 2)
        exp = torch.cat([torch.arange(r * 1000, (r + 1) * 1000, dtype=torch.float32) + 10000 * s for s in range(n)])
        assert torch.equal(z, exp)
        assert torch.equal(g, torch.cat([torch.full((777,), float(s)) for s in range(n)]))
        assert torch.all(xr == float(n))
        assert torch.allclose(xa, torch.full((64,), (n - 1) / 2))
        assert st["ops"] == 6 and st["segments"] > st["ops"]


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_worker_tasks():
    dev = torch.device("cuda", 0)
    a = torch.randn(1 << 20, device=dev)
    b = torch.randn(1 << 20, device=dev)
    c = torch.zeros(1 << 20, device=dev)
    h = torch.randn(4099, device=dev).to(torch.bfloat16)
    h2 = torch.randn(4099, device=dev).to(torch.bfloat16)
    hout = torch.zeros(4099, device=dev, dtype=torch.bfloat16)
    odd = torch.zeros(1001, device=dev, dtype=torch.uint8)
    odd_src = torch.arange(1001, device=dev).to(torch.uint8)
    flag = torch.zeros(2, device=dev, dtype=torch.int64)
    torch.cuda.synchronize(dev)
    w = uk.Worker(device=0, nlanes=2, idle_us=2000)
    w.copy(0, c.data_ptr(), a.data_ptr(), a.numel() * 4)
    w.reduce(0, c.data_ptr(), c.data_ptr(), b.data_ptr(), a.numel() * 4, torch.float32, "sum")
    w.signal(0, flag.data_ptr(), 3)
    w.wait_value(1, flag.data_ptr(), 3)  # lane 1 runs after lane 0's copy+reduce
    w.reduce(1, hout.data_ptr(), h.data_ptr(), h2.data_ptr(), 4099 * 2, torch.bfloat16, "max")
    w.copy(1, odd.data_ptr() + 1, odd_src.data_ptr() + 1, 999)  # unaligned byte copy
    w.wait_all()
    # a device-wide synchronisation must not hang on the idle worker: it quits after idle_us ...
    torch.cuda.synchronize(dev)
    first = w.kernel_launches
    assert first >= 1
    # ... and is relaunched transparently by the next task
    c2 = torch.zeros_like(c)
    torch.cuda.synchronize(dev)
    t = w.copy(1, c2.data_ptr(), c.data_ptr(), c.numel() * 4)
    w.wait(1, t)
    assert w.kernel_launches > first
    w.stop()
    assert w.error == 0
    torch.cuda.synchronize(dev)
    assert torch.equal(c2, c)
    assert torch.equal(c, a + b)
    assert torch.equal(hout, torch.maximum(h, h2))
    assert torch.equal(odd[1:1000], odd_src[1:1000]) and odd[0] == 0 and odd[1000] == 0
    assert flag[0].item() == 3


@pytest.mark.gpu
@pytest.mark.parametrize("n", [2, 4])
def test_gpu_ukcomm_collectives(n):
    comms = get_world(n)
    bufs = []
    for c in comms:  # candidates for the zero-copy path: allocated before the heaps are skewed
        with torch.cuda.device(c.device):
            b = c.empty(40000, dtype=torch.bfloat16)
            b.copy_(torch.full((40000,), float(c.rank + 1)))
            bufs.append(b)
    skew = comms[1]._c.alloc(54321, 256)  # rank 1's regions land at other offsets than everybody else's
    uks = [uk.UkCommunicator(c, nlanes=2, tile_bytes=64 << 10, staging_bytes=1 << 20) for c in comms]
    try:
        count = (1 << 19) + 12  # 2 MiB + tail: two staged segments
        ins = [torch.randn(count, generator=torch.Generator().manual_seed(r)) for r in range(n)]
        ref = torch.stack(ins).sum(0)
        xs, works = [], []
        for c, u in zip(comms, uks):
            with torch.cuda.device(c.device):
                s = torch.cuda.Stream(device=c.device)
                base = ins[c.rank].to(c.device)
                warm = base * 1.0  # load the elementwise kernel before any worker spins (lazy module loading)
                xs.append([s, base, None, None])
            torch.cuda.synchronize(c.device)
        for c, u, st in zip(comms, uks, xs):
            with torch.cuda.device(c.device), torch.cuda.stream(st[0]):
                x = st[1] * 1.0  # produced on the side stream, right before the collective
                works.append(u.all_reduce(x, "sum"))
                y = x * 2.0  # consumer on the same stream: ordered after the worker by cuStreamWaitValue64
                st[2], st[3] = x, y
        for (s, _, x, y), w in zip(xs, works):
            s.synchronize()
            w.wait()
            assert torch.allclose(x.cpu(), ref, rtol=1e-5, atol=1e-4)
            assert torch.allclose(y.cpu(), 2 * ref, rtol=1e-5, atol=2e-4)
        # zero-copy on symmetric buffers, ring algorithm, bf16
        for c in comms:
            torch.cuda.synchronize(c.device)
        works = []
        # zero-copy is only legal when the buffers sit at the same heap offset on every rank
        symmetric = len({c._c.heap_offset(b.data_ptr()) for c, b in zip(comms, bufs)}) == 1
        # one stream per rank: virtual ranks share a device, and a shared stream would serialise
        # rank 1's "inputs ready" write behind rank 0's wait for the collective (= deadlock)
        streams = [st[0] for st in xs]
        for c, u, b, s in zip(comms, uks, bufs, streams):
            with torch.cuda.device(c.device), torch.cuda.stream(s):
                works.append(u.all_reduce(b, "sum", algo="ring", symmetric=symmetric))
        for w in works:
            w.wait()
        for c, b in zip(comms, bufs):
            torch.cuda.synchronize(c.device)
            assert torch.all(b.float().cpu() == n * (n + 1) / 2)
        assert uks[0].stats()["zero_copy_ops"] == (1 if symmetric else 0)
        # all_to_all + all_gather + barrier
        per = 3001
        outs, works = [], []
        for c in comms:
            with torch.cuda.device(c.device):
                y = (torch.arange(n * per, dtype=torch.float32) + 1e5 * c.rank).to(c.device)
                z = torch.zeros(n * per, device=c.device)
                g = torch.zeros(n * 500, device=c.device)
                gi = torch.full((500,), float(c.rank)).to(c.device)
                outs.append((y, z, g, gi))
            torch.cuda.synchronize(c.device)
        for c, u, (y, z, g, gi), s in zip(comms, uks, outs, streams):
            with torch.cuda.device(c.device), torch.cuda.stream(s):
                works.append(u.all_to_all_single(z, y))
                works.append(u.all_gather_into_tensor(g, gi))
                works.append(u.barrier())
        for w in works:
            w.wait()
        for r, (y, z, g, gi) in enumerate(outs):
            torch.cuda.synchronize(comms[r].device)
            exp = torch.cat([torch.arange(r * per, (r + 1) * per, dtype=torch.float32) + 1e5 * s for s in range(n)])
            assert torch.equal(z.cpu(), exp)
            assert torch.equal(g.cpu(), torch.cat([torch.full((500,), float(s)) for s in range(n)]))
        del bufs
    finally:
        for u in uks:
            u.stop()
        comms[1]._c.free(skew)


def _mp_uk_worker(rank, n, uid, q):
    import torch

    from uccl_b200 import Communicator
    from uccl_b200 import ukernel as uk

    c = Communicator.init(uid, rank, n, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20, timeout_ms=20000)
    if rank == 1:
        c._c.alloc(4096, 256)  # asymmetric heaps across processes
    pg = uk.ProcessGroup(c, nlanes=2, tile_bytes=8192, staging_bytes=128 << 10)
    x = torch.arange(30000, dtype=torch.float32) + rank
    pg.all_reduce(x, "sum", )
    z = torch.zeros(n * 100, dtype=torch.int64)
    pg.all_to_all_single(z, torch.arange(n * 100, dtype=torch.int64) + 1000 * rank)
    pg.barrier()
    pg.shutdown()
    exp = sum(torch.arange(30000, dtype=torch.float32) + r for r in range(n))
    exp_z = torch.cat([torch.arange(rank * 100, (rank + 1) * 100, dtype=torch.int64) + 1000 * s for s in range(n)])
    q.put((rank, bool(torch.equal(x, exp)), bool(torch.equal(z, exp_z))))


def test_ukcomm_two_processes_over_shared_memory():
    """The FIFO workers of two real processes signal each other through the shm-backed symmetric heap."""
    import multiprocessing as mp

    ctx = mp.get_context("spawn")
    uid = Communicator.create_unique_id()
    q = ctx.Queue()
    ps = [ctx.Process(target=_mp_uk_worker, args=(r, 2, uid, q)) for r in range(2)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=120) for _ in range(2))
    [p.join(30) for p in ps]
    assert got == [(0, True, True), (1, True, True)]


def test_host_worker_wait_timeout_sets_error_word():
    """A WAIT task whose signal never arrives ends with an error word the producer sees, not a hang."""
    w = uk.Worker(device=-1, nlanes=1, timeout_ms=200)
    flag = torch.zeros(1, dtype=torch.int64)
    t = w.wait_value(0, flag.data_ptr(), 1)
    with pytest.raises(RuntimeError) as ei:
        w.wait(0, t, timeout_s=5.0)
    assert "error" in str(ei.value)
    assert w.error != 0


def test_reduce_scatter_and_broadcast_plans():
    for n in (1, 2, 3, 8):
        for lanes in (1, 3):
            assert uk.validate("reduce_scatter", 5008, n, lanes, 1024, 4) == ""
            for root in (0, n - 1):
                assert uk.validate("broadcast", 5000, n, lanes, 1024, root=root) == ""
    n, per = 4, 3000
    ins = [torch.randn(n * per, generator=torch.Generator().manual_seed(r)) for r in range(n)]
    outs = [torch.zeros(per) for _ in range(n)]
    assert uk.simulate("reduce_scatter", ins, outs, "sum", nlanes=2, tile_bytes=2048) == ""  # several tiles per lane
    ref = torch.stack(ins).sum(0).view(n, per)
    for r in range(n):
        assert torch.allclose(outs[r], ref[r], atol=1e-5)
    bins = [torch.full((4097,), float(r)) for r in range(n)]
    bouts = [torch.zeros(4097) for _ in range(n)]
    assert uk.simulate("broadcast", bins, bouts, nlanes=3, tile_bytes=1024, root=2) == ""
    for o in bouts:
        assert torch.equal(o, bins[2])


def test_host_ukcomm_reduce_scatter_broadcast():
    n = 4
    comms = Communicator.local_world(n, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20)

    def fn(c):
        pg = uk.ProcessGroup(c, nlanes=2, tile_bytes=4096, staging_bytes=64 << 10)
        x = (torch.arange(n * 7001, dtype=torch.float32) % 13) + c.rank   # 7001 per rank: odd size, several segments
        out = torch.zeros(7001)
        pg.reduce_scatter_tensor(out, x, "sum")
        avg = torch.zeros(16)
        pg.reduce_scatter_tensor(avg, torch.full((n * 16,), float(c.rank)), "avg")
        b = torch.full((70001,), float(c.rank), dtype=torch.bfloat16)
        pg.broadcast(b, src=3)
        pg.shutdown()
        return out, avg, b

    for r, (out, avg, b) in enumerate(_run_threads(comms, fn)):
        exp = sum((torch.arange(n * 7001, dtype=torch.float32) % 13) + s for s in range(n)).view(n, 7001)[r]
        assert torch.equal(out, exp)
        assert torch.allclose(avg, torch.full((16,), (n - 1) / 2))
        assert bool((b == 3).all())


# ------------------------------------------------------------------ user-authored programs (ukernel.dsl)
@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("nlanes", [1, 3])
def test_dsl_recursive_doubling_allreduce_simulated(n, nlanes):
    from uccl_b200.ukernel import dsl

    numel = 1000 + 8 * n
    prog = dsl.recursive_doubling_allreduce(n, numel * 4, elem_size=4, nlanes=nlanes)
    prog.validate()
    g = torch.Generator().manual_seed(n)
    ins = [torch.randn(numel, generator=g) for _ in range(n)]
    keep = [t.clone() for t in ins]
    outs = [torch.zeros(numel) for _ in range(n)]
    prog.simulate(ins, outs, "sum")
    exp = torch.stack(keep).sum(0)
    for r in range(n):
        assert torch.allclose(outs[r], exp, rtol=1e-5, atol=1e-5) and torch.equal(ins[r], keep[r])
    # the same program through its JSON form, in place (Out is In), with another operator
    again = dsl.Program.from_json(prog.to_json())
    assert again.to_dict() == prog.to_dict() and again.num_ops() == prog.num_ops()
    bufs = [t.clone() for t in keep]
    again.simulate(bufs, bufs, "max")
    for r in range(n):
        assert torch.equal(bufs[r], torch.stack(keep).max(0).values)


@pytest.mark.parametrize("n,root", [(2, 1), (5, 3), (8, 0)])
def test_dsl_binomial_broadcast_simulated(n, root):
    from uccl_b200.ukernel import dsl

    nbytes = 4096 + 16
    prog = dsl.binomial_broadcast(n, nbytes, root=root, nlanes=2)
    prog.validate()
    ins = [torch.full((nbytes,), r, dtype=torch.uint8) for r in range(n)]
    outs = [torch.zeros(nbytes, dtype=torch.uint8) for _ in range(n)]
    prog.simulate(ins, outs)
    for r in range(n):
        assert bool((outs[r] == root).all())
    # log2(n) rounds: the root sends ceil(log2 n) times per lane, not n - 1 times
    sends_root = sum(1 for o in prog.ops[root] if o["kind"] == "send")
    assert sends_root == 2 * (n - 1).bit_length()


def test_dsl_rejects_broken_programs():
    from uccl_b200.ukernel import dsl
    from uccl_b200.ukernel.dsl import In, Out, Program, Scratch

    p = Program("unmatched", 2, 1, 64, 64, 64)
    p.isend(0, 1, Scratch(0), In(0), 64)  # never waited for
    with pytest.raises(ValueError, match="match|never"):
        p.validate()
    p = Program("overflow", 2, 1, 64, 64, 64)
    p.send(0, 1, Scratch(32), In(0), 64)  # 32 + 64 > 64 bytes of scratch
    with pytest.raises(ValueError, match="outside"):
        p.validate()
    p = Program("misaligned", 2, 1, 64, 64, 0)
    p.copy(0, Out(8), In(0), 16)
    with pytest.raises(ValueError, match="misaligned"):
        p.validate()
    p = Program("lanes", 2, 1, 64, 64, 0)
    p.copy(0, Out(0), In(0), 16, lane=1)
    with pytest.raises(ValueError):
        p.validate()
    p = Program("reduce-size", 2, 1, 64, 64, 0, elem_size=4)
    p.reduce(0, Out(0), In(0), In(16), 6)
    with pytest.raises(ValueError):
        p.validate()
    with pytest.raises(ValueError):
        dsl.recursive_doubling_allreduce(6, 1024)
    with pytest.raises(ValueError):
        Program.from_json('{"format": "something else"}')
    # a receive that can never be satisfied in order: rank 1 waits for rank 0 before posting what rank 0 waits for
    p = Program("deadlock", 2, 1, 64, 64, 64)
    h01 = p.isend(0, 1, Scratch(0), In(0), 16)
    p.wait(h01)
    p.send(1, 0, Scratch(0), In(0), 16)
    p.validate()  # written in a valid order: fine
    ins = [torch.zeros(64, dtype=torch.uint8) for _ in range(2)]
    p.simulate(ins, [t.clone() for t in ins])


@pytest.mark.parametrize("n", [2, 4])
def test_dsl_program_on_host_ukcomm(n):
    """The same programs executed rank by rank on the ukernel communicator (host backend: FIFOs drained by threads),
    staged (ordinary tensors, in place and out of place) and after a built-in collective on the same lanes."""
    from uccl_b200.ukernel import dsl

    comms = Communicator.local_world(n, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20)
    numel = 3000
    rd = dsl.recursive_doubling_allreduce(n, numel * 4, elem_size=4, nlanes=2)
    bc = dsl.binomial_broadcast(n, numel * 4, root=n - 1, nlanes=2)
    text = rd.to_json()

    def fn(c):
        u = uk.UkCommunicator(c, nlanes=2, tile_bytes=4096, staging_bytes=64 << 10)
        warm = torch.full((100,), float(c.rank))
        u.all_reduce(warm, "sum").wait()
        x = torch.arange(numel, dtype=torch.float32) * (c.rank + 1)
        y = torch.zeros(numel)
        dsl.Program.from_json(text).run(u, x, y, "sum").wait()
        z = x.clone()
        rd.run(u, z, op="max").wait()  # in place
        b = torch.full((numel,), float(c.rank))
        bc.run(u, b).wait()
        with pytest.raises(ValueError):
            dsl.recursive_doubling_allreduce(2 * n, numel * 4).run(u, x, y)
        st = u.stats()
        u.stop()
        return x, y, z, b, warm, st

    for r, (x, y, z, b, warm, st) in enumerate(_run_threads(comms, fn)):
        base = torch.arange(numel, dtype=torch.float32)
        assert torch.equal(x, base * (r + 1))
        assert torch.allclose(y, base * (n * (n + 1) / 2))
        assert torch.equal(z, base * n)
        assert bool((b == float(n - 1)).all()) and bool((warm == n * (n - 1) / 2).all())
        assert st["ops"] == 4


def test_custom_collective_example_cpu():
    """examples/custom_collective.py --cpu: a hand-written two-level all-reduce, validated, simulated, shipped as JSON
    and executed by four host-backend ranks."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "custom_collective.py"), "--cpu", "--ranks", "4",
                        "--numel", "5000"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "pair_reduce_doubling_4" in r.stdout and "executed on the host backend: ok" in r.stdout


def test_planner_cpp_unit_under_sanitizers(tmp_path):
    """tests/cpp/uk_plan_test.cc: planner + validator + bounds checker + simulator as a plain C++ program under
    ASan + UBSan (no Python, no CUDA) -- the reference's C++ unit layer for the ukernel CCL."""
    import os
    import shutil
    import subprocess

    if shutil.which("g++") is None:
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    csrc = os.path.join(root, "uccl_b200", "csrc")
    exe = str(tmp_path / "uk_plan_test")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-I" + csrc, "-I/usr/local/cuda/include",
           os.path.join(root, "tests/cpp/uk_plan_test.cc"), os.path.join(csrc, "ukernel/uk_plan.cc"),
           os.path.join(csrc, "coll/host_coll.cc"), "-o", exe, "-lpthread"]
    b = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=dict(os.environ, ASAN_OPTIONS="detect_leaks=0"))
    assert r.returncode == 0 and "PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("n", [2, 3])
def test_process_group_surface_of_the_reference(n):
    """ReduceOp / Work / properties / uneven all_to_all_single / functional API of the reference's ukernel_ccl
    (experimental/ukernel/py/ukernel_ccl/__init__.py:19-25,171-430) on the host backend."""
    comms = Communicator.local_world(n, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20)

    def fn(c):
        pg = uk.ProcessGroup(c, nlanes=2, tile_bytes=4096, staging_bytes=64 << 10)
        r = c.rank
        assert (pg.rank, pg.world_size, pg.gpu_id, pg.backend) == (r, n, -1, "ukernel")
        assert pg.same_host((r + 1) % n) and pg.peer_transport(r) == "self" and pg.peer_transport((r + 1) % n) == "host-shm"
        x = torch.full((300,), float(r + 1))
        assert pg.all_reduce(x, uk.ReduceOp.MAX, tile_bytes=64 << 10, num_flows=2) is None
        w = pg.all_reduce(torch.ones(8), op=uk.ReduceOp.SUM, async_op=True)
        assert isinstance(w, uk.Work)
        w.wait()
        assert w.is_completed()
        with pytest.raises(ValueError):
            pg.all_reduce(x, uk.ReduceOp.BAND)
        # uneven splits: rank r sends (d + 1) rows of 3 values to rank d
        isp = [d + 1 for d in range(n)]
        osp = [r + 1] * n
        inp = torch.cat([torch.full((d + 1, 3), float(10 * r + d)) for d in range(n)])
        out = torch.zeros(sum(osp), 3)
        pg.all_to_all_single(out, inp, output_split_sizes=osp, input_split_sizes=isp)
        y = torch.arange(n * 4, dtype=torch.float32) + 100 * r
        z = torch.zeros(n * 4)
        pg.all_to_all_single(z, y, output_split_sizes=[4] * n, input_split_sizes=[4] * n)  # explicit but even
        with pytest.raises(ValueError):
            pg.all_to_all_single(out, inp, output_split_sizes=osp, input_split_sizes=[1] * n)
        pg.barrier()
        pg.shutdown()
        return x, out, z

    for r, (x, out, z) in enumerate(_run_threads(comms, fn)):
        assert bool((x == float(n)).all())
        exp = torch.cat([torch.full((r + 1, 3), float(10 * s + r)) for s in range(n)])
        assert torch.equal(out, exp)
        assert torch.equal(z, torch.cat([torch.arange(r * 4, r * 4 + 4, dtype=torch.float32) + 100 * s for s in range(n)]))


def test_functional_api_on_a_default_group():
    c = Communicator.local_world(1, host=True, heap_bytes=160 << 20, stage_bytes=1 << 20)[0]
    assert not uk.is_initialized()
    with pytest.raises(RuntimeError):
        uk.get_rank()
    pg = uk.init_process_group("ukernel", comm=c, nlanes=1, tile_bytes=4096, staging_bytes=64 << 10)
    try:
        assert uk.is_initialized() and uk.get_rank() == 0 and uk.get_world_size() == 1 and uk.get_rank(pg) == 0
        t = torch.full((10,), 2.0)
        uk.all_reduce(t, uk.ReduceOp.SUM)
        o = torch.zeros(6)
        uk.all_to_all_single(o, torch.arange(6, dtype=torch.float32))
        uk.barrier()
        assert bool((t == 2.0).all()) and torch.equal(o, torch.arange(6, dtype=torch.float32))
        with pytest.raises(RuntimeError):
            uk.init_process_group("ukernel", comm=c)
        with pytest.raises(ValueError):
            uk.destroy_process_group()
            uk.init_process_group("nccl", comm=c)
    finally:
        uk.destroy_process_group()
    assert not uk.is_initialized()


def test_rank_addressed_p2p_communicator():
    """`uccl_b200.ukernel.p2p.Communicator` -- the reference's ukernel_p2p surface (experimental/ukernel/py/
    ukernel_p2p.cpp:407-444): connect / accept by rank through the exchanger, published buffer ids, isend / irecv with
    byte offsets, request polling, named barriers.  Three ranks in one process, host memory."""
    import socket
    import threading

    from uccl_b200 import compat
    from uccl_b200.ukernel.p2p import Communicator

    compat.install_ukernel()
    import ukernel_p2p

    assert ukernel_p2p.Communicator is Communicator
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    W = 3
    res, comms, errs = {}, {}, []

    def body(r):
        c = ukernel_p2p.Communicator(gpu_id=-1, rank=r, world_size=W, exchanger_ip="127.0.0.1", exchanger_port=port,
                                     transport="auto")
        comms[r] = c
        assert c.rank == r and c.world_size == W
        for p in range(W):
            if p != r:
                assert c.accept_peer(p) if r < p else c.connect_peer(p)
        nxt, prv = (r + 1) % W, (r - 1) % W
        assert c.same_host(nxt) and c.peer_transport(nxt) == "tcp"
        buf = torch.zeros(64, dtype=torch.float32)
        with pytest.raises(ValueError):
            c.reg_rdma(0, buf)
        assert c.reg_rdma(100 + r, buf, publish=True)
        assert c.wait_mr(nxt, 100 + nxt) and not c.wait_mr(nxt, 999, timeout_ms=50)
        src = torch.arange(64, dtype=torch.float32) + 1000 * r
        rq = c.irecv(prv, buf, offset=64, len=128)
        sq = c.isend(nxt, src, offset=32, len=128, remote_buffer_id=100 + nxt, remote_offset=64)
        assert rq and sq and c.wait_finish_multi([sq, rq]) and c.poll(sq)
        with pytest.raises(ValueError):
            c.isend(nxt, src, offset=200, len=128)
        with pytest.raises(RuntimeError):
            c.isend(nxt, src, remote_buffer_id=4242)
        res[r] = buf.clone()
        assert c.barrier() and c.barrier("phase2", 5000)
        done, got = torch.full((1,), float(r)), torch.zeros(1)
        if r % 2 == 0:
            c.send(nxt, done)
            c.recv(prv, got)
        else:
            c.recv(prv, got)
            c.send(nxt, done)
        assert float(got) == float(prv)
        assert c.unreg_rdma(100 + r) and not c.unreg_rdma(100 + r)
        assert c.barrier("end")

    def run(r):
        try:
            body(r)
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(W)]
    [t.start() for t in ths]
    [t.join(120) for t in ths]
    assert not errs, errs
    assert not any(t.is_alive() for t in ths), "a rank is stuck"
    for r in range(W):
        prv = (r - 1) % W
        exp = torch.zeros(64)
        exp[16:48] = torch.arange(8, 40, dtype=torch.float32) + 1000 * prv
        assert torch.equal(res[r], exp)
    for r in sorted(comms, reverse=True):
        comms[r].close()
    compat.uninstall()
    with pytest.raises(ValueError):
        Communicator(gpu_id=-1, rank=0, world_size=1, exchanger_port=port, transport="carrier-pigeon")


def test_functional_api_standalone_world():
    """`ukernel_ccl.init_process_group(rank=, world_size=, gpu_id=, exchanger_ip=, exchanger_port=, transport=)` as in
    the reference's test_collective.py: no torch.distributed, the world is rendezvoused through the exchanger that
    rank 0 serves.  Two processes, host memory (tests/uk_ccl_worker.py)."""
    import os
    import socket
    import subprocess
    import sys

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", EXCHANGER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(root, "tests", "uk_ccl_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=180)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        outs.append(o)
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"rank {r} ok" in o, o[-3000:]
