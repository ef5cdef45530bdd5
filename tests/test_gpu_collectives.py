"""GPU numerics tests of the native collective kernels against plain fp32/fp64 PyTorch
references.  Ranks are real GPUs when the box has enough of them, otherwise *virtual ranks*
(one stream per rank on device 0; peers' heaps are then other allocations of the same GPU, the
kernel code path -- flags, barriers, LL packets, slicing -- is identical)."""
import pytest
import torch

from helpers import get_world, run_ranks

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _ref(ins, op):
    ref = torch.stack([x.double() for x in ins])
    if op == "sum":
        return ref.sum(0)
    if op == "avg":
        return ref.sum(0) / len(ins)
    if op == "max":
        return ref.max(0).values
    if op == "min":
        return ref.min(0).values
    return ref.prod(0)


def _inputs(n, count, dtype, seed=0):
    g = torch.Generator().manual_seed(seed)
    if dtype.is_floating_point:
        return [(torch.randn(count, generator=g) * 2).to(dtype) for _ in range(n)]
    return [torch.randint(-50, 50, (count,), generator=g).to(dtype) for _ in range(n)]


def _tol(dtype):
    if dtype == torch.bfloat16:
        return dict(rtol=2e-2, atol=6e-2)
    if dtype == torch.float16:
        return dict(rtol=4e-3, atol=1e-2)
    if dtype == torch.float32:
        return dict(rtol=1e-5, atol=1e-5)
    return dict(rtol=0, atol=0)


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("count", [1, 7, 1024, 4099, 65536 + 3])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_allreduce_oneshot(n, count, dtype):
    comms = get_world(n)
    if count * dtype.itemsize > (256 << 10):
        pytest.skip("beyond LL range")
    ins = _inputs(n, count, dtype)
    exp = _ref(ins, "sum")

    def prepare(c):
        return ins[c.rank].to(c.device)

    def launch(c, x):
        c.all_reduce(x, "sum", algo="oneshot_ll")

    outs = run_ranks(comms, prepare, launch)
    for o in outs:
        assert torch.allclose(o.cpu().double(), exp, **_tol(dtype))
        assert torch.equal(o.cpu(), outs[0].cpu())  # bitwise identical across ranks


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("dtype,op", [(torch.float32, "sum"), (torch.bfloat16, "avg"), (torch.float16, "max"),
                                      (torch.int32, "sum"), (torch.int64, "min"), (torch.float64, "sum"),
                                      (torch.uint8, "max")])
def test_allreduce_large(n, sym, dtype, op):
    comms = get_world(n)
    count = (3 << 20) // dtype.itemsize + 5  # > stage chunk for the staged path, odd tail
    ins = _inputs(n, count, dtype)
    if dtype == torch.uint8:
        ins = [x.abs() for x in ins]
    exp = _ref(ins, op)
    if op == "avg" and not dtype.is_floating_point:
        exp = torch.div(exp * n, n, rounding_mode="trunc")

    def prepare(c):
        if sym:
            x = c.empty(count, dtype=dtype)
            x.copy_(ins[c.rank])
            return x
        return ins[c.rank].to(c.device)

    def launch(c, x):
        c.all_reduce(x, op, algo="twoshot_p2p" if sym else "staged_p2p")

    outs = run_ranks(comms, prepare, launch)
    for o in outs:
        assert torch.allclose(o.cpu().double(), exp.to(dtype).double(), **_tol(dtype))
        assert torch.equal(o.cpu(), outs[0].cpu())


@pytest.mark.parametrize("n", [2, 8])
def test_allreduce_auto_sizes(n):
    """AUTO selection across the size range incl. repeated launches (epoch/parity reuse)."""
    comms = get_world(n)
    for count in [3, 300, 40000, 100000, 1 << 20]:
        for rep in range(3):
            ins = _inputs(n, count, torch.float32, seed=rep)
            exp = _ref(ins, "sum")
            outs = run_ranks(comms, lambda c: ins[c.rank].to(c.device), lambda c, x: c.all_reduce(x, "sum"))
            for o in outs:
                assert torch.allclose(o.cpu().double(), exp, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("n", [2, 4])
def test_allreduce_fused_scale_and_cast(n):
    comms = get_world(n)
    count = 1 << 18
    ins = _inputs(n, count, torch.float32)
    exp = (_ref(ins, "sum") * 0.125).to(torch.bfloat16).double()

    def prepare(c):
        x = c.empty(count, dtype=torch.float32)
        x.copy_(ins[c.rank])
        return x, c.empty(count, dtype=torch.bfloat16)

    def launch(c, st):
        c.all_reduce(st[0], "sum", out=st[1], scale=0.125)

    outs = run_ranks(comms, prepare, launch)
    for _, o in outs:
        assert torch.allclose(o.cpu().double(), exp, rtol=2e-2, atol=2e-2)
    # staged variant (plain tensors)
    outs = run_ranks(comms, lambda c: (ins[c.rank].to(c.device), torch.empty(count, dtype=torch.bfloat16, device=c.device)),
                     lambda c, st: c.all_reduce(st[0], "sum", out=st[1], scale=0.125))
    for _, o in outs:
        assert torch.allclose(o.cpu().double(), exp, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("sym_out,sym_in", [(True, False), (False, True), (False, False)])
def test_allgather(n, sym_out, sym_in):
    comms = get_world(n)
    for count in [5, 4096, (1 << 20) + 16]:
        ins = _inputs(n, count, torch.float32, seed=count)
        exp = torch.cat(ins)

        def prepare(c):
            x = c.empty(count, dtype=torch.float32) if sym_in else torch.empty(count, device=c.device)
            x.copy_(ins[c.rank])
            out = c.empty(n * count, dtype=torch.float32) if sym_out else torch.empty(n * count, device=c.device)
            out.zero_()
            return x, out

        outs = run_ranks(comms, prepare, lambda c, st: c.all_gather(st[1], st[0]))
        for _, o in outs:
            assert torch.equal(o.cpu(), exp)


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("sym", [True, False])
@pytest.mark.parametrize("dtype,op", [(torch.float32, "sum"), (torch.bfloat16, "sum"), (torch.int32, "max")])
def test_reduce_scatter(n, sym, dtype, op):
    comms = get_world(n)
    for count in [8, 1000, (1 << 18) + 4]:
        ins = _inputs(n, n * count, dtype, seed=count)
        exp = _ref(ins, op).to(dtype).view(n, count)

        def prepare(c):
            x = c.empty(n * count, dtype=dtype) if sym else torch.empty(n * count, dtype=dtype, device=c.device)
            x.copy_(ins[c.rank])
            return x, torch.zeros(count, dtype=dtype, device=c.device)

        outs = run_ranks(comms, prepare, lambda c, st: c.reduce_scatter(st[1], st[0], op))
        for r, (_, o) in enumerate(outs):
            assert torch.allclose(o.cpu().double(), exp[r].double(), **_tol(dtype))


@pytest.mark.parametrize("n", [2, 4, 8])
@pytest.mark.parametrize("ll", [True, False])
def test_ll_exchange_vs_barrier_paths(n, ll):
    """all_gather / all_to_all / reduce_scatter through the barrier-free LL-packet kernels (ll=True) and
    through the barrier-based kernels (ll=False) on the same inputs; repeated calls flip the packet parity."""
    comms = get_world(n)
    for c in comms:
        c.set_xchg_ll_max((1 << 20) if ll else -1)
    try:
        for it, count in enumerate([4, 1024, 20480, 4, 65536]):
            for dtype, op in ((torch.float32, "sum"), (torch.bfloat16, "max"), (torch.int32, "sum")):
                if (count * dtype.itemsize) % 16:
                    continue
                ins = _inputs(n, n * count, dtype, seed=100 * it + count)
                # all_gather of every rank's first `count` elements (out-of-place, then in-place)
                exp_ag = torch.cat([x[:count] for x in ins])

                def prep_ag(c):
                    return ins[c.rank][:count].to(c.device), torch.zeros(n * count, dtype=dtype, device=c.device)

                outs = run_ranks(comms, prep_ag, lambda c, st: c.all_gather(st[1], st[0]))
                for _, o in outs:
                    assert torch.equal(o.cpu(), exp_ag)

                def prep_ag_inplace(c):
                    out = torch.zeros(n * count, dtype=dtype, device=c.device)
                    out[c.rank * count:(c.rank + 1) * count].copy_(ins[c.rank][:count])
                    return out

                outs = run_ranks(comms, prep_ag_inplace,
                                 lambda c, o: c.all_gather(o, o[c.rank * count:(c.rank + 1) * count]))
                for o in outs:
                    assert torch.equal(o.cpu(), exp_ag)
                # all_to_all
                outs = run_ranks(comms, lambda c: (ins[c.rank].to(c.device), torch.zeros(n * count, dtype=dtype, device=c.device)),
                                 lambda c, st: c.all_to_all(st[1], st[0]))
                for r, (_, o) in enumerate(outs):
                    exp = torch.cat([ins[s][r * count:(r + 1) * count] for s in range(n)])
                    assert torch.equal(o.cpu(), exp)
                # reduce_scatter
                exp_rs = _ref(ins, op).view(n, count)
                outs = run_ranks(comms, lambda c: (ins[c.rank].to(c.device), torch.zeros(count, dtype=dtype, device=c.device)),
                                 lambda c, st: c.reduce_scatter(st[1], st[0], op))
                for r, (_, o) in enumerate(outs):
                    assert torch.allclose(o.cpu().double(), exp_rs[r], **_tol(dtype))
    finally:
        for c in comms:
            c.set_xchg_ll_max(0)


@pytest.mark.parametrize("n", [2, 4, 8])
def test_reduce_scatter_push_staging(n):
    """Plain-buffer reduce_scatter with the push staging variant, several double-buffered chunks
    (1 MiB stage), odd sizes, and the same results as the pull variant bit for bit."""
    comms = get_world(n, heap_mb=160, stage_mb=1, max_ctas=4)
    for c in comms:
        c.set_xchg_ll_max(-1)
    try:
        for dtype, op, count in ((torch.float32, "sum", (1 << 18) + 4), (torch.bfloat16, "max", 70001),
                                 (torch.int32, "sum", 1000)):
            ins = _inputs(n, n * count, dtype, seed=count)
            exp = _ref(ins, op).view(n, count)
            res = {}
            for push in (False, True):
                for c in comms:
                    c.set_rs_push(push)
                outs = run_ranks(comms, lambda c: (ins[c.rank].to(c.device), torch.zeros(count, dtype=dtype, device=c.device)),
                                 lambda c, st: c.reduce_scatter(st[1], st[0], op))
                res[push] = [o.cpu() for _, o in outs]
                for r in range(n):
                    got = res[push][r].double()
                    bad = (~torch.isclose(got, exp[r], **_tol(dtype))).nonzero().flatten()
                    assert bad.numel() == 0, (f"push={push} rank={r} {dtype} {op} count={count}: {bad.numel()} bad, "
                                              f"first at {bad[:4].tolist()}, last at {bad[-1].item()}, "
                                              f"got {got[bad[:3]].tolist()} want {exp[r][bad[:3]].tolist()}")
            for r in range(n):
                if comms[0].has_multicast and dtype.is_floating_point:
                    # the pull variant reduces inside the switch (multimem.ld_reduce): same value up to rounding
                    assert torch.allclose(res[True][r].double(), res[False][r].double(), **_tol(dtype))
                else:
                    assert torch.equal(res[True][r], res[False][r])
    finally:
        for c in comms:
            c.set_rs_push(True)
            c.set_xchg_ll_max(0)


@pytest.mark.parametrize("n", [2, 8])
def test_broadcast_reduce_alltoall(n):
    comms = get_world(n)
    count = 50001
    ins = _inputs(n, count, torch.float32)
    root = n - 1
    # broadcast (plain and symmetric outputs)
    for sym in (False, True):
        def prepare(c):
            x = c.empty(count, dtype=torch.float32) if sym else torch.empty(count, device=c.device)
            x.copy_(ins[c.rank])
            return x
        outs = run_ranks(comms, prepare, lambda c, x: c.broadcast(x, root=root))
        for o in outs:
            assert torch.equal(o.cpu(), ins[root])
    # reduce
    exp = _ref(ins, "sum")
    outs = run_ranks(comms, lambda c: ins[c.rank].to(c.device), lambda c, x: c.reduce(x, root=0, op="sum"))
    assert torch.allclose(outs[0].cpu().double(), exp, rtol=1e-5, atol=1e-5)
    # all_to_all: equal splits
    per = 1000
    a_ins = [torch.arange(n * per, dtype=torch.float32) + 10000 * r for r in range(n)]
    for sym_out, sym_in in ((True, False), (False, True), (False, False)):
        def prepare(c):
            x = c.empty(n * per, dtype=torch.float32) if sym_in else torch.empty(n * per, device=c.device)
            x.copy_(a_ins[c.rank])
            o = c.empty(n * per, dtype=torch.float32) if sym_out else torch.empty(n * per, device=c.device)
            o.zero_()
            return x, o
        outs = run_ranks(comms, prepare, lambda c, st: c.all_to_all(st[1], st[0]))
        for r, (_, o) in enumerate(outs):
            exp_a = torch.cat([a_ins[s][r * per:(r + 1) * per] for s in range(n)])
            got = o.cpu()
            bad = (got != exp_a).nonzero().flatten()
            assert bad.numel() == 0, (f"all_to_all sym_out={sym_out} sym_in={sym_in} rank={r}: {bad.numel()} mismatches, "
                                      f"first at {int(bad[0])}: got {float(got[bad[0]])} want {float(exp_a[bad[0]])}")


@pytest.mark.parametrize("n", [2, 4])
def test_alltoallv(n):
    comms = get_world(n)
    # rank s sends (s + d + 1) * 10 elements to rank d
    cnt = [[(s + d + 1) * 10 for d in range(n)] for s in range(n)]
    ins = [torch.cat([torch.full((cnt[s][d],), float(100 * s + d)) for d in range(n)]) for s in range(n)]

    def prepare(c):
        r = c.rank
        x = ins[r].to(c.device)
        out = torch.zeros(sum(cnt[s][r] for s in range(n)), device=c.device)
        return x, out

    def launch(c, st):
        r = c.rank
        c.all_to_all_v(st[1], st[0], cnt[r], [cnt[s][r] for s in range(n)])

    outs = run_ranks(comms, prepare, launch)
    for r, (_, o) in enumerate(outs):
        exp = torch.cat([torch.full((cnt[s][r],), float(100 * s + r)) for s in range(n)])
        assert torch.equal(o.cpu(), exp)


def test_nvls_paths_if_available():
    """Real multi-GPU boxes only: exercise multimem.ld_reduce/st kernels."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(torch.cuda.device_count(), 8)
    n = 8 if n >= 8 else (4 if n >= 4 else 2)
    comms = get_world(n)
    if not comms[0].has_multicast:
        pytest.skip("no NVLS multicast on this box")
    for dtype in (torch.float32, torch.bfloat16):
        for count in (1000, 1 << 20):
            ins = _inputs(n, count, dtype)
            exp = _ref(ins, "sum")

            def prepare(c):
                x = c.empty(count, dtype=dtype)
                x.copy_(ins[c.rank])
                return x

            for algo in ("oneshot_mc", "twoshot_nvls"):
                if algo == "oneshot_mc" and count * dtype.itemsize > (256 << 10):
                    continue
                outs = run_ranks(comms, prepare, lambda c, x: c.all_reduce(x, "sum", algo=algo))
                for o in outs:
                    assert torch.allclose(o.cpu().double(), exp, **_tol(dtype))
            outs = run_ranks(comms, lambda c: ins[c.rank].to(c.device),
                             lambda c, x: c.all_reduce(x, "sum", algo="staged_nvls"))
            for o in outs:
                assert torch.allclose(o.cpu().double(), exp, **_tol(dtype))


def test_cuda_graph_replay():
    """Device-side epochs make every kernel CUDA-graph replayable (no host-side launch counters):
    capture one allreduce per rank, replay three times with new inputs."""
    n = 2
    comms = get_world(n)
    count = 1 << 16
    xs = [c.empty(count, dtype=torch.float32) for c in comms]
    small = [c.empty(256, dtype=torch.float32) for c in comms]
    for x, s in zip(xs, small):
        x.zero_()
        s.zero_()
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=c.device) for c in comms]
    graphs = []
    for c, st, x, s in zip(comms, streams, xs, small):
        with torch.cuda.device(c.device):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                c.all_reduce(x, "sum", algo="twoshot_p2p")
                c.all_reduce(s, "sum", algo="oneshot_ll")
            graphs.append(g)
    for rep in range(3):
        for r, (x, s) in enumerate(zip(xs, small)):
            x.fill_(float(rep + r + 1))
            s.fill_(float(10 * rep + r))
        torch.cuda.synchronize()
        for c, st, g in zip(comms, streams, graphs):
            with torch.cuda.device(c.device), torch.cuda.stream(st):
                g.replay()
        for st in streams:
            st.synchronize()
        exp = sum(float(rep + r + 1) for r in range(n))
        exp_s = sum(float(10 * rep + r) for r in range(n))
        for x, s in zip(xs, small):
            assert bool((x == exp).all()), (rep, x[:4])
            assert bool((s == exp_s).all()), (rep, s[:4])


def test_device_trace():
    n = 2
    comms = get_world(n)
    for c in comms:
        c.enable_trace(4096)
    xs = [c.empty(1 << 16, dtype=torch.float32) for c in comms]
    for x in xs:
        x.fill_(1.0)
    run_ranks(comms, lambda c: xs[c.rank], lambda c, x: c.all_reduce(x, "sum", algo="twoshot_p2p"))
    for c in comms:
        ev = c.dump_trace()
        kinds = [e["event"] for e in ev]
        assert kinds.count("barrier_enter") == kinds.count("barrier_exit") >= 2
        assert "kernel_begin" in kinds
        assert all(ev[i]["t_ns"] <= ev[i + 1]["t_ns"] for i in range(len(ev) - 1))
        c.disable_trace()
    assert bool((xs[0] == 2.0).all())


def test_torch_mem_pool_in_symmetric_heap():
    """Tensors created under comm.use_mem_pool() live in the symmetric heap: collectives take the
    zero-copy algorithms on them, although the ranks allocate at different offsets."""
    n = 2
    comms = get_world(n)
    count = 1 << 20
    ins = _inputs(n, count, torch.float32, seed=5)
    exp = _ref(ins, "sum")
    keep = []

    def prepare(c):
        with c.use_mem_pool():
            if c.rank == 1:
                keep.append(torch.empty(12345, device=c.device))  # skew rank 1's offsets
            x = torch.empty(count, device=c.device)
        assert c._c.in_heap(x.data_ptr(), x.numel() * 4)
        x.copy_(ins[c.rank])
        algo, _ = c.select_allreduce(count * 4, True, torch.float32)
        assert algo.startswith("twoshot")
        return x

    outs = run_ranks(comms, prepare, lambda c, x: c.all_reduce(x, "sum"))
    for o in outs:
        assert torch.allclose(o.cpu().double(), exp, rtol=1e-5, atol=1e-5)
    from uccl_b200 import _native

    st = _native.C().pool_stats()
    assert st["allocs"] >= 3 and st["fallback_allocs"] == 0
    y = torch.empty(8, device=comms[0].device)  # outside the context: ordinary allocator again
    assert not comms[0]._c.in_heap(y.data_ptr(), 32)
