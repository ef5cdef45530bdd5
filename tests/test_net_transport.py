"""The scale-out path on a CPU-only box: multipath reliable datagram transport (csrc/net), its NCCL net
plugin, the rank-group collectives on top, and the hierarchical multi-node communicator -- all over
loopback UDP, with injected packet loss (the reference tests its transports with two hosts; the protocol
logic is the same on one).  SURVEY N1 / N2 / N4 / N6 / N7."""
import os
import subprocess
import sys
import threading

import pytest
import torch

from uccl_b200 import Communicator, net
from uccl_b200.parallel import MultiNodeCommunicator

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.timeout(300)


@pytest.fixture(scope="module")
def engine_test_exe(tmp_path_factory):
    exe = tmp_path_factory.mktemp("net") / "net_engine_test"
    csrc = os.path.join(ROOT, "uccl_b200", "csrc")
    subprocess.run(["g++", "-std=c++17", "-O2", "-I" + csrc, os.path.join(ROOT, "tests/cpp/net_engine_test.cc"),
                    os.path.join(csrc, "net/net_engine.cc"), "-lpthread", "-o", str(exe)], check=True)
    return exe


@pytest.mark.parametrize("drop,cc,isn", [(0, "swift", None), (5, "swift", None), (3, "none", None), (2, "timely", None),
                                         (2, "eqds", None), (3, "swift", 4294967000)])
def test_engine_loopback_with_loss(engine_test_exe, drop, cc, isn):
    """C++ level: handshake, eager/unexpected + rendezvous messages, bidirectional transfers, 64 pipelined
    messages, spraying over every path, retransmissions under loss, dead-peer abort; the last case pins the
    initial sequence numbers just below 2^32 so that every window computation crosses the wrap."""
    env = dict(os.environ)
    if isn is not None:
        env["UCCL_B200_NET_ISN"] = str(isn)
    r = subprocess.run([str(engine_test_exe), str(drop), cc, "8"], capture_output=True, text=True, timeout=240, env=env)
    sys.stdout.write(r.stdout)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PASS" in r.stdout


def test_nccl_net_plugin_vtable(tmp_path):
    """dlopen the plugin and drive the ncclNet v8 vtable like NCCL's proxy thread."""
    from uccl_b200 import _build

    _build.build()
    plugin = _build.nccl_net_plugin_path()
    assert plugin.exists()
    syms = subprocess.run(["nm", "-D", str(plugin)], capture_output=True, text=True).stdout
    assert "ncclNetPlugin_v8" in syms
    assert " U cuda" not in syms and " U cu" not in syms  # host-only object: loads into any NCCL process
    exe = tmp_path / "net_plugin_test"
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests/cpp/net_plugin_test.cc"), "-ldl", "-lpthread",
                    "-o", str(exe)], check=True)
    env = dict(os.environ, UCCL_B200_NET_IFNAME="lo")
    r = subprocess.run([str(exe), str(plugin)], capture_output=True, text=True, timeout=200, env=env)
    sys.stdout.write(r.stdout + r.stderr)
    assert r.returncode == 0 and "PASS" in r.stdout
    assert "dev0 lo" in r.stdout


def test_python_engine_tensors_and_stats():
    a = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.03)
    b = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.03)
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("fb", b.accept(lid)))
    t.start()
    fa = a.connect("127.0.0.1", b.port, lid)
    t.join()
    fb = box["fb"]
    x = torch.randn(3_000_000)
    y = torch.zeros_like(x)
    small = torch.arange(100, dtype=torch.int32)
    got_small = torch.zeros(100, dtype=torch.int32)
    w1 = b.irecv(fb, y)
    w2 = a.isend(fa, x)
    a.send(fa, small)  # eager, queued behind the large message
    assert w1.wait(60000) == x.numel() * 4 and w2.wait(60000) == x.numel() * 4
    assert b.recv(fb, got_small, 60000) == 400
    assert torch.equal(x, y) and torch.equal(small, got_small)
    # a receive that is too small fails loudly instead of truncating
    big = torch.ones(5000)
    tiny = torch.zeros(10)
    wr = b.irecv(fb, tiny)
    a.send(fa, big, 60000)
    with pytest.raises(RuntimeError, match="larger than the posted receive"):
        wr.wait(60000)
    st = a.flow_stats(fa)
    assert st["state"] == 2  # FL_ESTABLISHED
    assert all(p > 0 for p in st["path_tx"]) and len(st["path_tx"]) == 4
    assert st["fast_rexmit"] + st["rto_rexmit"] > 0 and st["srtt_us"] > 0
    assert a.stats()["dropped_tx"] > 0
    with pytest.raises(ValueError):
        a.isend(fa, torch.zeros(4, 4).t())
    a.close(fa)


class _Exchange:
    """In-process all-gather of python objects between rank threads (stands in for a TCPStore)."""

    def __init__(self, n):
        self.slots, self.bar = [None] * n, threading.Barrier(n)

    def for_rank(self, r):
        def ex(obj):
            self.slots[r] = obj
            self.bar.wait()
            out = list(self.slots)
            self.bar.wait()
            return out

        return ex


def _run_threads(n, fn):
    out, errs = [None] * n, []

    def body(r):
        try:
            out[r] = fn(r)
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=body, args=(r,)) for r in range(n)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    return out


@pytest.mark.parametrize("n", [2, 3, 4])
def test_net_communicator_collectives(n):
    ex = _Exchange(n)
    g = torch.Generator().manual_seed(n)
    ins = [torch.randint(-50, 50, (10007,), generator=g).float() for _ in range(n)]

    def fn(r):
        c = net.NetCommunicator(r, n, ex.for_rank(r), engine=net.Engine(bind_ip="127.0.0.1", paths=3, drop_prob=0.01))
        res = {}
        x = ins[r].clone()
        c.all_reduce(x)
        res["sum"] = x
        res["max"] = c.all_reduce(ins[r].clone(), "max")
        res["avg"] = c.all_reduce(ins[r].clone(), "avg")
        res["tiny"] = c.all_reduce(torch.tensor([float(r + 1)]))  # fewer elements than ranks
        ag = torch.zeros(n * 33)
        c.all_gather(ag, torch.full((33,), float(r)))
        res["ag"] = ag
        rs = torch.zeros(50)
        c.reduce_scatter(rs, torch.arange(n * 50, dtype=torch.float32) * (r + 1))
        res["rs"] = rs
        b = torch.full((1000,), float(r))
        c.broadcast(b, root=n - 1)
        res["b"] = b
        a2a = torch.zeros(n * 4, dtype=torch.int64)
        c.all_to_all(a2a, torch.arange(n * 4, dtype=torch.int64) + 100 * r)
        res["a2a"] = a2a
        c.barrier()
        res["stats"] = c.stats()
        c.close()
        return res

    outs = _run_threads(n, fn)
    ref = torch.stack(ins)
    for r, o in enumerate(outs):
        assert torch.equal(o["sum"], ref.sum(0)) and torch.equal(o["max"], ref.max(0).values)
        assert torch.allclose(o["avg"], ref.sum(0) / n)
        assert o["tiny"].item() == n * (n + 1) / 2
        assert torch.equal(o["ag"].view(n, 33)[:, 0], torch.arange(n, dtype=torch.float32))
        assert torch.equal(o["rs"], torch.arange(n * 50, dtype=torch.float32).view(n, 50)[r] * (n * (n + 1) / 2))
        assert torch.equal(o["b"], torch.full((1000,), float(n - 1)))
        assert torch.equal(o["a2a"], torch.cat([torch.arange(4) + 4 * r + 100 * s for s in range(n)]))
        assert o["stats"]["engine"]["tx_pkts"] > 0


def test_multinode_communicator_two_nodes_of_two():
    """2 'nodes' x 2 local ranks in one process: host symmetric heaps inside a node, datagram transport
    between nodes.  Every hierarchical collective must equal the flat 4-rank reference."""
    N, L = 2, 2
    W = N * L
    nodes = [Communicator.local_world(L, host=True, heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=30000)
             for _ in range(N)]
    rails = [_Exchange(N) for _ in range(L)]
    g = torch.Generator().manual_seed(9)
    ins = [torch.randint(-9, 9, (4099,), generator=g).float() for _ in range(W)]

    def fn(gr):
        k, l = divmod(gr, L)
        nc = net.NetCommunicator(k, N, rails[l].for_rank(k), engine=net.Engine(bind_ip="127.0.0.1", paths=2, drop_prob=0.01))
        m = MultiNodeCommunicator(nodes[k][l], nc)
        assert (m.rank, m.world_size, m.node_rank, m.local_rank) == (gr, W, k, l)
        res = {}
        res["sum"] = m.all_reduce(ins[gr].clone())             # 4099 is not a multiple of L: padded path
        res["avg"] = m.all_reduce(ins[gr].clone(), "avg")
        res["max"] = m.all_reduce(ins[gr][:4096].clone(), "max")
        o16 = torch.zeros(4099, dtype=torch.bfloat16)
        m.all_reduce(ins[gr].clone(), "sum", out=o16, scale=0.5)  # the DDP compress-hook contract
        res["scaled"] = o16
        m.pipeline_bytes = 1024                                  # 4099 floats -> 2050 per shard -> 9 chunks in flight
        res["pipe_sum"] = m.all_reduce(ins[gr].clone())
        res["pipe_avg"] = m.all_reduce(ins[gr].clone(), "avg")
        m.pipeline_bytes = 8 << 20
        ag = torch.zeros(W * 17)
        m.all_gather(ag, torch.full((17,), float(gr)))
        res["ag"] = ag
        rs = torch.zeros(25)
        m.reduce_scatter(rs, torch.arange(W * 25, dtype=torch.float32) * (gr + 1))
        res["rs"] = rs
        for root in (0, 3):
            b = torch.full((1001,), float(gr))
            m.broadcast(b, root=root)
            res[f"b{root}"] = b
        a2a = torch.zeros(W * 3)
        m.all_to_all(a2a, torch.arange(W * 3, dtype=torch.float32) + 100 * gr)
        res["a2a"] = a2a
        mate, rail = gr ^ 1, (gr + L) % W                      # box-mate and rail-mate in one group
        r1, r2 = torch.zeros(3000), torch.zeros(3000)
        m.batch_send_recv([("recv", r1, mate), ("recv", r2, rail), ("send", torch.full((3000,), 10.0 + gr), mate),
                           ("send", torch.full((3000,), 20.0 + gr), rail)])
        res["p2p"] = (bool((r1 == 10.0 + mate).all()), bool((r2 == 20.0 + rail).all()))
        m.barrier()
        m.close()
        return res

    outs = _run_threads(W, fn)
    ref = torch.stack(ins)
    for gr, o in enumerate(outs):
        assert torch.equal(o["sum"], ref.sum(0))
        assert torch.allclose(o["avg"], ref.sum(0) / W)
        assert torch.equal(o["max"], ref[:, :4096].max(0).values)
        assert torch.equal(o["scaled"], (ref.sum(0) * 0.5).to(torch.bfloat16))
        assert torch.equal(o["pipe_sum"], ref.sum(0)) and torch.allclose(o["pipe_avg"], ref.sum(0) / W)
        assert torch.equal(o["ag"].view(W, 17)[:, 0], torch.arange(W, dtype=torch.float32))
        assert torch.equal(o["rs"], torch.arange(W * 25, dtype=torch.float32).view(W, 25)[gr] * (W * (W + 1) / 2))
        assert torch.equal(o["b0"], torch.zeros(1001)) and torch.equal(o["b3"], torch.full((1001,), 3.0))
        assert torch.equal(o["a2a"], torch.cat([torch.arange(3, dtype=torch.float32) + 3 * gr + 100 * s for s in range(W)]))
        assert o["p2p"] == (True, True)


_MP_SCRIPT = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["UB_ROOT"])
from uccl_b200 import net
from uccl_b200.parallel import MultiNodeCommunicator
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
m = MultiNodeCommunicator.from_torch_dist(local_size=2, engine=net.Engine(bind_ip="127.0.0.1", paths=2), host=True,
                                         heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=60000)
x = torch.arange(100000, dtype=torch.float32) + rank
m.all_reduce(x)
exp = torch.arange(100000, dtype=torch.float32) * world + sum(range(world))
ok = bool(torch.equal(x, exp))
m.barrier()
print(f"rank {rank} node {m.node_rank} local {m.local_rank} ok={ok}", flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


def test_multinode_from_torch_dist_four_processes(tmp_path):
    """The documented bootstrap: torchrun world (gloo) -> node groups (shared-memory heaps) + rail groups
    (datagram flows) -> hierarchical all-reduce, 4 real processes."""
    script = tmp_path / "mn.py"
    script.write_text(_MP_SCRIPT)
    env = dict(os.environ, UB_ROOT=ROOT, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=4",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=240, env=env)
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok=True") == 4


def _pg_worker(rank, world, path, q):
    import torch.distributed as dist

    os.environ["UCCL_B200_LOCAL_SIZE"] = "2"       # 4 ranks = 2 "boxes" of 2
    os.environ["UCCL_B200_NET_BIND_IP"] = "127.0.0.1"
    os.environ["UCCL_B200_NET_PATHS"] = "2"
    import uccl_b200.parallel.pg  # noqa: F401  (registers the backend)
    from uccl_b200.parallel import MultiNodeCommunicator

    torch.set_num_threads(1)
    dist.init_process_group("uccl_b200", rank=rank, world_size=world, store=dist.FileStore(path, world))
    ok = []
    x = torch.arange(5001, dtype=torch.float32) + rank
    dist.all_reduce(x)
    ok.append(torch.equal(x, torch.arange(5001, dtype=torch.float32) * world + sum(range(world))))
    g = torch.empty(world * 4, dtype=torch.int64)
    dist.all_gather_into_tensor(g, torch.full((4,), rank, dtype=torch.int64))
    ok.append(g.view(world, 4)[:, 0].tolist() == list(range(world)))
    b = torch.full((9,), float(rank))
    dist.broadcast(b, src=2)
    ok.append(bool((b == 2.0).all()))
    rs = torch.empty(10)
    dist.reduce_scatter_tensor(rs, torch.arange(world * 10, dtype=torch.float32) * (rank + 1))
    ok.append(torch.equal(rs, torch.arange(world * 10, dtype=torch.float32).view(world, 10)[rank] * sum(range(1, world + 1))))
    red = torch.full((6,), float(rank + 1))
    dist.reduce(red, dst=1)
    ok.append(bool((red == (10.0 if rank == 1 else rank + 1)).all()))
    a_out = torch.empty(world * 3)
    dist.all_to_all_single(a_out, torch.arange(world * 3, dtype=torch.float32) + 10 * rank)
    ok.append(a_out.tolist() == [float(10 * s + 3 * rank + i) for s in range(world) for i in range(3)])
    # unequal splits: rank s sends (s + d) % 3 + 1 elements to rank d
    sc = [(rank + d) % 3 + 1 for d in range(world)]
    rc = [(s + rank) % 3 + 1 for s in range(world)]
    v_in = torch.cat([torch.full((sc[d],), float(100 * rank + d)) for d in range(world)])
    v_out = torch.empty(sum(rc))
    dist.all_to_all_single(v_out, v_in, output_split_sizes=rc, input_split_sizes=sc)
    ok.append(torch.equal(v_out, torch.cat([torch.full((rc[s],), float(100 * s + rank)) for s in range(world)])))
    # DDP across the two boxes
    torch.manual_seed(0)
    model = torch.nn.Linear(8, 4)
    ddp = torch.nn.parallel.DistributedDataParallel(model)
    ddp(torch.full((2, 8), float(rank + 1))).sum().backward()
    ok.append(torch.allclose(model.weight.grad, torch.full((4, 8), sum(2.0 * (r + 1) for r in range(world)) / world)))
    # point to point: inside the box and along the rail
    for peer in (rank ^ 1, (rank + 2) % world):
        s, r_ = torch.full((5,), float(rank)), torch.empty(5)
        if rank < peer:
            dist.send(s, peer)
            dist.recv(r_, peer)
        else:
            dist.recv(r_, peer)
            dist.send(s, peer)
        ok.append(bool((r_ == peer).all()))
    dist.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_torch_backend_spans_two_boxes():
    """dist.init_process_group("uccl_b200") with LOCAL size 2 and world size 4: the backend builds a
    MultiNodeCommunicator (shm heap per box + datagram rails) and every c10d collective + DDP works."""
    import multiprocessing as mp
    import tempfile

    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        ps = [ctx.Process(target=_pg_worker, args=(r, world, os.path.join(d, "store"), q)) for r in range(world)]
        [p.start() for p in ps]
        got = sorted(q.get(timeout=200) for _ in range(world))
        [p.join(60) for p in ps]
    for rank, ok in got:
        assert all(ok), (rank, ok)


@pytest.mark.parametrize("staged", [False, True])
def test_p2p_net_channel_kv_blocks(staged):
    """Cross-box P2P: a list of KV blocks through NetChannel, in place for host tensors and through the
    chunked pinned-staging pipeline (the path GPU tensors take), with loss on the wire."""
    from uccl_b200.p2p import NetChannel

    ea = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.01)
    eb = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.01)
    srv = NetChannel.listen(eb, chunk_bytes=100_000, force_staging=staged)
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("ch", srv.accept()))
    t.start()
    cli = NetChannel.connect(ea, ("127.0.0.1", eb.port, srv.address[2]), chunk_bytes=100_000, force_staging=staged)
    t.join()
    g = torch.Generator().manual_seed(3)
    blocks = [torch.randn(n, generator=g).to(dt) for n, dt in ((250_001, torch.float32), (7, torch.bfloat16), (131072, torch.float16), (1, torch.int64))]
    outs = [torch.zeros_like(b) for b in blocks]
    rt = threading.Thread(target=lambda: box.setdefault("n", srv.recv_tensors(outs)))
    rt.start()
    sent = cli.send_tensors(blocks)
    rt.join()
    assert box["n"] == sent == sum(b.numel() * b.element_size() for b in blocks)
    for a, b in zip(blocks, outs):
        assert torch.equal(a, b)
    # a receiver that expects a different layout is told so instead of getting garbage
    et = threading.Thread(target=lambda: box.setdefault("err", _catch(lambda: srv.recv_tensors(outs[:2]))))
    et.start()
    with pytest.raises(RuntimeError, match="rejected"):
        cli.send_tensors(blocks[:1] + blocks[2:3])
    et.join()
    assert isinstance(box["err"], RuntimeError) and "announced" in str(box["err"])
    cli.close()
    srv.close()


def _catch(fn):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        return e
    return None


def test_expert_parallel_buffer_across_boxes():
    """DeepEP Buffer over a group that spans 2 boxes x 2 ranks: the token exchange is the hierarchical two-hop
    all-to-all (shared-memory heap inside a box, datagram rails between boxes).  Same oracle as the intranode
    tests: received rows in source-rank-major order, combine == x * fan-out; plus the low-latency pair."""
    from test_host_ep import _inputs
    from uccl_b200.ep import Buffer

    N, L = 2, 2
    W = N * L
    nodes = [Communicator.local_world(L, host=True, heap_bytes=128 << 20, stage_bytes=1 << 20, timeout_ms=30000)
             for _ in range(N)]
    rails = [_Exchange(N) for _ in range(L)]
    T, H, K, E, M = 37, 256, 3, 8, 48
    e_per = E // W
    xs, idxs, ws = _inputs(W, T, H, K, E, seed=5)

    def fn(gr):
        k, l = divmod(gr, L)
        nc = net.NetCommunicator(k, N, rails[l].for_rank(k), engine=net.Engine(bind_ip="127.0.0.1", paths=2, drop_prob=0.005))
        m = MultiNodeCommunicator(nodes[k][l], nc)
        b = Buffer(comm=m, num_nvl_bytes=1 << 20, num_rdma_bytes=1 << 20, low_latency_mode=True)
        assert (b.rank, b.group_size, b.get_num_rdma_ranks()) == (gr, W, N)
        tpr, _, tpe, inr, _ = b.get_dispatch_layout(idxs[gr], E)
        rx, ri, rw, pe, h, _ = b.internode_dispatch(xs[gr], None, tpr, None, inr, tpe, idxs[gr], ws[gr])
        comb, _, _ = b.internode_combine(rx, h, topk_weights=rw)
        lx, cnt, lh, _, _ = b.low_latency_dispatch(xs[gr], idxs[gr], M, E, use_fp8=False)
        lout, _, _ = b.low_latency_combine(lx, idxs[gr], ws[gr], lh)
        m.close()
        return dict(rx=rx, inr=inr, comb=comb, pe=pe, cnt=cnt, lout=lout)

    outs = _run_threads(W, fn)
    for r, o in enumerate(outs):
        exp = torch.cat([xs[s][outs[s]["inr"][:, r].nonzero().flatten()] for s in range(W)])
        assert torch.equal(o["rx"], exp)
        assert torch.allclose(o["comb"].float(), xs[r].float() * o["inr"].sum(1).float()[:, None], rtol=2e-2, atol=1e-1)
        counts = [int(sum((idxs[s] == r * e_per + e).sum() for s in range(W))) for e in range(e_per)]
        assert o["pe"] == counts and o["cnt"].tolist() == counts
        wsum = torch.where(idxs[r] >= 0, ws[r], torch.zeros_like(ws[r])).sum(1)
        assert torch.allclose(o["lout"].float(), xs[r].float() * wsum[:, None], rtol=3e-2, atol=1e-1)


def test_black_holed_path_is_quarantined():
    """One of four paths silently drops everything (a dead ECMP route): transfers keep completing, the loss
    streak quarantines the path, and it carries only a sliver of the traffic."""
    a = net.Engine(bind_ip="127.0.0.1", paths=4)
    b = net.Engine(bind_ip="127.0.0.1", paths=4)
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("fb", b.accept(lid)))
    t.start()
    fa = a.connect("127.0.0.1", b.port, lid)
    t.join()
    a.set_path_drop(1, 1.0)
    x = torch.randn(4_000_000)
    y = torch.zeros_like(x)
    for _ in range(2):
        w = b.irecv(box["fb"], y)
        a.send(fa, x, 120000)
        w.wait(120000)
        assert torch.equal(x, y)
    st = a.flow_stats(fa)
    assert st["path_bans"] >= 1
    healthy = [st["path_tx"][i] for i in (0, 2, 3)]
    assert st["path_tx"][1] < 0.5 * min(healthy), st["path_tx"]
    a.set_path_drop(-1, 0.0)


def _coll_mb_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), UCCL_B200_LOCAL_SIZE="2",
                      UCCL_B200_NET_BIND_IP="127.0.0.1", UCCL_B200_NET_PATHS="2")
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uccl_b200 import collective as C

    ctx = C.init_collective(local_gpu_idx=-1, heap_bytes=128 << 20)
    ok = [sorted(ctx.remote_peers) == [p for p in range(world) if p // 2 != rank // 2]]
    # ring step over every kind of hop (inside the box: P2P engine; between boxes: datagram channel)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    x = torch.full((300_000,), float(rank))
    y = torch.zeros(300_000)
    hs = C.batch_isend_irecv([ctx.P2POp("irecv", y, prv), ctx.P2POp("isend", x, nxt)])
    C.wait_all(hs)
    ok.append(bool((y == prv).all()))
    # two sends to one remote peer keep their order
    far = (rank + 2) % world
    a, b = torch.full((10,), 1.0 + rank), torch.full((10,), 100.0 + rank)
    ra, rb = torch.zeros(10), torch.zeros(10)
    hs = [C.irecv(ra, far), C.irecv(rb, far), C.isend(a, far), C.isend(b, far)]
    C.wait_all(hs)
    ok.append(bool((ra == 1.0 + far).all() and (rb == 100.0 + far).all()))
    g = torch.zeros(world * 50)
    C.allgather(torch.full((50,), float(rank)), g)
    ok.append(g.view(world, 50)[:, 0].tolist() == [float(r) for r in range(world)])
    t = torch.arange(1000, dtype=torch.float32) + rank
    C.all_reduce(t, "sum")
    ok.append(torch.equal(t, torch.arange(1000, dtype=torch.float32) * world + sum(range(world))))
    C.barrier()
    C.finalize_collective()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_collective_module_across_boxes():
    """`uccl_b200.collective` (the reference's `uccl.collective` surface) on 2 boxes x 2 ranks: send/recv pick
    the P2P engine inside a box and the datagram channel between boxes; native collectives go hierarchical."""
    import multiprocessing as mp
    import socket

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_coll_mb_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in ps]
    for rank, ok in got:
        assert all(ok), (rank, ok)


def test_many_flows_share_one_engine():
    """48 flows between two engines (NCCL opens several comms per peer and channel): interleaved transfers on
    all of them complete and stay separated."""
    a = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.005)
    b = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.005)
    lid = b.listen()
    nflows = 48
    fa = [a._native.connect_async("127.0.0.1", b.port, lid) for _ in range(nflows)]
    fb = [b.accept(lid) for _ in range(nflows)]
    import time

    t0 = time.time()
    while any(a.flow_state(f) != 2 for f in fa):
        assert time.time() - t0 < 30
        time.sleep(0.001)
    # accept order is not connect order: every flow introduces itself
    tags = [torch.tensor([i], dtype=torch.int64) for i in range(nflows)]
    ws = [a.isend(f, t) for f, t in zip(fa, tags)]
    ident = {}
    for f in fb:
        who = torch.zeros(1, dtype=torch.int64)
        b.recv(f, who, 30000)
        ident[int(who)] = f
    [w.wait(30000) for w in ws]
    assert sorted(ident) == list(range(nflows))
    xs = [torch.full((25_000,), float(i)) for i in range(nflows)]
    ys = [torch.zeros(25_000) for _ in range(nflows)]
    rws = [b.irecv(ident[i], ys[i]) for i in range(nflows)]
    sws = [a.isend(fa[i], xs[i]) for i in range(nflows)]
    for w in rws + sws:
        w.wait(60000)
    for i in range(nflows):
        assert bool((ys[i] == i).all())
    assert a.stats()["flows"] >= nflows


def test_gpu_nic_matching_is_rail_aligned():
    """HGX-like tree: two root complexes, two PCIe switches each, one NIC next to each GPU pair."""
    from uccl_b200.net.topology import nic_for_gpu, pci_distance

    def dev(root, sw, leaf):
        return f"/sys/devices/pci0000:{root}/0000:{root}:01.0/0000:{sw}:00.0/0000:{leaf}:00.0"

    gpus = {0: dev("17", "18", "1a"), 1: dev("17", "18", "1b"), 2: dev("17", "28", "2a"), 3: dev("17", "28", "2b"),
            4: dev("97", "98", "9a"), 5: dev("97", "98", "9b"), 6: dev("97", "a8", "aa"), 7: dev("97", "a8", "ab")}
    nics = {"mlx0": dev("17", "18", "1c"), "mlx1": dev("17", "28", "2c"), "mlx2": dev("97", "98", "9c"),
            "mlx3": dev("97", "a8", "ac"), "eno1": "/sys/devices/pci0000:00/0000:00:1f.6"}
    ifs = [(n, f"10.0.0.{i}") for i, n in enumerate(nics)]
    assert pci_distance(gpus[0], nics["mlx0"]) < pci_distance(gpus[0], nics["mlx1"]) < pci_distance(gpus[0], nics["mlx2"])
    picks = [nic_for_gpu(g, local_rank=g, interfaces=ifs, nic_path=nics.get, gpu_path=gpus.get)[0] for g in range(8)]
    assert picks == ["mlx0", "mlx0", "mlx1", "mlx1", "mlx2", "mlx2", "mlx3", "mlx3"]
    # no PCI information (VMs, this container): NICs are shared round robin by local rank
    rr = [nic_for_gpu(None, local_rank=l, interfaces=ifs[:4], nic_path=lambda n: None)[0] for l in range(6)]
    assert rr == ["mlx0", "mlx1", "mlx2", "mlx3", "mlx0", "mlx1"]
    assert nic_for_gpu(None, interfaces=[])[1] == "127.0.0.1"


def test_net_communicator_stripes_large_messages_over_engines():
    """Two engine threads per rank: messages above the stripe threshold are cut in two and both engines carry
    traffic; small messages stay on the primary engine; results are unchanged."""
    n = 2
    ex = _Exchange(n)
    ins = [torch.arange(600_000, dtype=torch.float32) * (r + 1) for r in range(n)]

    def fn(r):
        mk = lambda: net.Engine(bind_ip="127.0.0.1", paths=2)  # noqa: E731
        c = net.NetCommunicator(r, n, ex.for_rank(r), engine=mk(), extra_engines=[mk()], stripe_min_bytes=256 << 10)
        assert len(c.engines) == 2
        x = ins[r].clone()
        c.all_reduce(x)                      # 1.2 MB segments: striped
        small = c.all_reduce(torch.full((10,), float(r)))
        g = torch.zeros(n * 300_000)
        c.all_gather(g, torch.full((300_000,), float(r)))
        c.barrier()
        st = c.stats()
        c.close()
        return x, small, g, [e["tx_bytes"] for e in st["engines"]]

    outs = _run_threads(n, fn)
    for r, (x, small, g, txb) in enumerate(outs):
        assert torch.equal(x, ins[0] + ins[1]) and bool((small == 1.0).all())
        assert torch.equal(g.view(n, -1)[:, 0], torch.arange(n, dtype=torch.float32))
        assert min(txb) > 1_000_000, txb     # both engines moved megabytes


def test_engine_edge_cases():
    """Mismatched path counts negotiate the minimum; connecting to nobody times out with an error; requests
    that are still pending when the peer goes away fail instead of hanging."""
    a = net.Engine(bind_ip="127.0.0.1", paths=8, rto_min_us=2000, rto_abort=6)
    b = net.Engine(bind_ip="127.0.0.1", paths=3)
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("fb", b.accept(lid)))
    t.start()
    fa = a.connect("127.0.0.1", b.port, lid)
    t.join()
    x, y = torch.randn(500_000), torch.zeros(500_000)
    w = b.irecv(box["fb"], y)
    a.send(fa, x, 60000)
    w.wait(60000)
    assert torch.equal(x, y)
    used = [p for p in a.flow_stats(fa)["path_tx"] if p > 0]
    assert len(used) == 3                                   # only the paths both sides have
    # nobody listens on this port: the SYN is never answered
    s = __import__("socket").socket(__import__("socket").AF_INET, __import__("socket").SOCK_DGRAM)
    s.bind(("127.0.0.1", 0))
    dead_port = s.getsockname()[1]
    with pytest.raises(RuntimeError, match="timed out"):
        a.connect("127.0.0.1", dead_port, 1, timeout_ms=400)
    s.close()
    # the peer disappears while a rendezvous send is parked and a receive is posted
    big = torch.ones(1_000_000)
    ws = a.isend(fa, big)                                   # waits for b's RTR, which never comes
    wr = a.irecv(fa, torch.zeros(10))
    del b, w                                                # b's engine shuts down: FIN reaches a
    import gc

    gc.collect()
    with pytest.raises(RuntimeError, match="peer closed"):
        wr.wait(30000)
    with pytest.raises(RuntimeError):                       # keep-alive probe -> retransmission limit -> flow error
        ws.wait(60000)
    assert a.flow_state(fa) == 5


@pytest.mark.parametrize("n,algo", [(2, "auto"), (3, "ring"), (4, "fullmesh"), (4, "ring")])
def test_ukernel_plans_over_the_datagram_transport(n, algo):
    """The ukernel planner's tile DAGs executed by the message-transport adapter (Send = header + payload message,
    placement on arrival, Recv = arrival counter), with loss on the wire; several tiles and lanes per collective."""
    from uccl_b200 import ukernel

    ex = _Exchange(n)
    g = torch.Generator().manual_seed(7 * n)
    ins = [torch.randint(-20, 20, (50_003,), generator=g).float() for _ in range(n)]

    def fn(r):
        u = ukernel.UkNetCommunicator(r, n, ex.for_rank(r), engine=net.Engine(bind_ip="127.0.0.1", paths=2, drop_prob=0.005),
                                      nlanes=2, tile_bytes=16 << 10)
        res = {}
        res["sum"] = u.all_reduce(ins[r].clone(), algo=algo)
        o = torch.zeros_like(ins[r])
        u.all_reduce(ins[r], "max", out=o, algo=algo)            # out of place
        res["max"] = o
        a2a = torch.zeros(n * 3000, dtype=torch.int64)
        u.all_to_all_single(a2a, torch.arange(n * 3000, dtype=torch.int64) + 100_000 * r)
        res["a2a"] = a2a
        ag = torch.zeros(n * 5000)
        u.all_gather_into_tensor(ag, torch.full((5000,), float(r)))
        res["ag"] = ag
        rs = torch.zeros(4000)
        u.reduce_scatter_tensor(rs, torch.arange(n * 4000, dtype=torch.float32) * (r + 1))
        res["rs"] = rs
        b = torch.full((30_000,), float(r))
        u.broadcast(b, root=n - 1)
        res["b"] = b
        u.barrier()
        res["stats"] = u.stats()
        return res

    outs = _run_threads(n, fn)
    ref = torch.stack(ins)
    for r, o in enumerate(outs):
        assert torch.equal(o["sum"], ref.sum(0)) and torch.equal(o["max"], ref.max(0).values)
        assert torch.equal(o["a2a"], torch.cat([torch.arange(3000) + 3000 * r + 100_000 * s for s in range(n)]))
        assert torch.equal(o["ag"].view(n, -1)[:, 0], torch.arange(n, dtype=torch.float32))
        assert torch.equal(o["rs"], torch.arange(n * 4000, dtype=torch.float32).view(n, 4000)[r] * (n * (n + 1) / 2))
        assert bool((o["b"] == n - 1).all())
        assert o["stats"]["sends"] > 0 and o["stats"]["recvs"] > 0


def test_proxy_link_put_with_signal_across_boxes():
    """The network half of the EP CPU proxy: box s writes a pattern into the heap of (box d, local rank l) and then
    bumps a counter there; whenever the counter shows k completed puts, the k blocks are already in place
    (ordering of put-then-add on one flow), with loss on the wire."""
    from uccl_b200.ep.proxy import ProxyLink

    nb, L, HB = 3, 2, 2 << 20
    ex = _Exchange(nb)
    heaps = [[torch.zeros(HB, dtype=torch.uint8) for _ in range(L)] for _ in range(nb)]
    blocks, BS = 24, 20_000
    barrier = threading.Barrier(nb)

    def fn(b):
        rail = net.NetCommunicator(b, nb, ex.for_rank(b), engine=net.Engine(bind_ip="127.0.0.1", paths=2, drop_prob=0.01))
        link = ProxyLink(rail, heaps[b])
        ok = True
        for d in range(nb):
            if d == b:
                continue
            for k in range(blocks):
                l = k % L
                # block k of sender b lands at a per-sender region of local rank l on box d
                off = 4096 + (b * blocks + k) * BS
                link.put(d, l, off, torch.full((BS,), (17 * b + k) % 251 + 1, dtype=torch.uint8))
                link.add(d, l, 8 * b, 1)                      # counter of sender b at offset 8*b of that heap
        # consumer side: poll my counters; every time one advances, the announced blocks must be complete
        import time

        t0 = time.time()
        seen = {(s, l): 0 for s in range(nb) if s != b for l in range(L)}
        while any(v < blocks // L for v in seen.values()):
            assert time.time() - t0 < 60, seen
            for (s, l), have in list(seen.items()):
                cnt = int(heaps[b][l][8 * s: 8 * s + 8].view(torch.int64).item())
                for j in range(have, cnt):
                    k = j * L + l                                # the j-th block sender s aimed at local rank l
                    off = 4096 + (s * blocks + k) * BS
                    ok &= bool((heaps[b][l][off: off + BS] == (17 * s + k) % 251 + 1).all())
                seen[(s, l)] = cnt
            time.sleep(0.0005)
        link.flush()
        barrier.wait()
        st = link.stats()
        rail.close()
        return ok, st

    outs = _run_threads(nb, fn)
    for ok, st in outs:
        assert ok
        assert st["puts"] == (nb - 1) * blocks and st["applied_writes"] == (nb - 1) * blocks
        assert st["applied_adds"] == (nb - 1) * blocks and st["bytes_in"] == (nb - 1) * blocks * BS


def _free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _native_mn_worker(rank, world, uid, q):
    os.environ.update(UCCL_B200_NET_BIND_IP="127.0.0.1", UCCL_B200_NET_PATHS="2", UCCL_B200_MN_PIPELINE_BYTES="32768")
    torch.set_num_threads(1)
    from uccl_b200.parallel import NativeMultiNodeCommunicator

    m = NativeMultiNodeCommunicator.init(uid, rank, world, local_size=2, host=True, heap_bytes=192 << 20, stage_bytes=1 << 20,
                                         timeout_ms=30000)
    ok = [(m.rank, m.world_size, m.node_rank, m.local_rank, m.num_nodes) == (rank, world, rank // 2, rank % 2, 2)]
    g = torch.Generator().manual_seed(1234)
    ins = [torch.randn(70_001, generator=g) for _ in range(world)]          # same on every rank
    for dt, tol in ((torch.float32, 1e-5), (torch.bfloat16, 3e-2), (torch.float16, 2e-3), (torch.float64, 1e-12)):
        x = ins[rank].clone().to(dt)
        m.all_reduce(x)                                                        # pipelined path (70001 elements)
        ref = torch.stack([t.to(dt).double() for t in ins]).sum(0)
        ok.append(torch.allclose(x.double(), ref, rtol=tol, atol=tol * 8))
    xi = torch.arange(5000, dtype=torch.int32) * (rank + 1)
    m.all_reduce(xi, "max")
    ok.append(torch.equal(xi, torch.arange(5000, dtype=torch.int32) * world))
    avg = ins[rank].clone()
    m.all_reduce(avg, "avg")
    ok.append(torch.allclose(avg, torch.stack(ins).mean(0), rtol=1e-5, atol=1e-6))
    out16 = torch.zeros(70_001, dtype=torch.bfloat16)
    m.all_reduce(ins[rank].clone(), "sum", out=out16, scale=0.25)
    ok.append(torch.allclose(out16.float(), torch.stack(ins).sum(0) * 0.25, rtol=2e-2, atol=2e-2))
    ag = torch.zeros(world * 333)
    m.all_gather(ag, torch.full((333,), float(rank)))
    ok.append(ag.view(world, 333)[:, 0].tolist() == [float(r) for r in range(world)])
    rs = torch.zeros(100)
    m.reduce_scatter(rs, torch.arange(world * 100, dtype=torch.float32) * (rank + 1), "sum")
    ok.append(torch.equal(rs, torch.arange(world * 100, dtype=torch.float32).view(world, 100)[rank] * 10))
    b = torch.full((12345,), float(rank))
    m.broadcast(b, root=3)
    ok.append(bool((b == 3).all()))
    a2a = torch.zeros(world * 7)
    m.all_to_all(a2a, torch.arange(world * 7, dtype=torch.float32) + 100 * rank)
    ok.append(a2a.tolist() == [float(100 * s + 7 * rank + i) for s in range(world) for i in range(7)])
    sc = [(rank + 2 * d) % 4 + 1 for d in range(world)]
    rc = [(s_ + 2 * rank) % 4 + 1 for s_ in range(world)]
    vout = torch.zeros(sum(rc))
    m.all_to_all_v(vout, torch.cat([torch.full((sc[d],), float(100 * rank + d)) for d in range(world)]), sc, rc)
    ok.append(torch.equal(vout, torch.cat([torch.full((rc[s_],), float(100 * s_ + rank)) for s_ in range(world)])))
    r1 = torch.zeros(2000)
    m.batch_send_recv([("recv", r1, (rank + 2) % world), ("send", torch.full((2000,), float(rank)), (rank + 2) % world)])
    ok.append(bool((r1 == (rank + 2) % world).all()))
    m.barrier()
    q.put((rank, ok))


def test_native_multinode_communicator_four_processes():
    """The C++ MultiComm through its Python wrapper: 2 boxes x 2 processes on the host backend, all dtypes of the
    rail reduction against torch references, pipelined all-reduce, fused scale + cast, every collective."""
    import multiprocessing as mp

    world = 4
    ctx = mp.get_context("spawn")
    uid = Communicator.create_unique_id()
    q = ctx.Queue()
    ps = [ctx.Process(target=_native_mn_worker, args=(r, world, uid, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in ps]
    for rank, ok in got:
        assert all(ok), (rank, ok)


def test_multinode_error_paths():
    """Things that must fail loudly: point-to-point across rails of different boxes, scale on integer tensors,
    CUDA tensors handed to the raw engine, non-contiguous tensors."""
    N, L = 2, 2
    W = N * L
    nodes = [Communicator.local_world(L, host=True, heap_bytes=96 << 20, stage_bytes=1 << 20, timeout_ms=30000)
             for _ in range(N)]
    rails = [_Exchange(N) for _ in range(L)]

    def fn(gr):
        k, l = divmod(gr, L)
        nc = net.NetCommunicator(k, N, rails[l].for_rank(k), engine=net.Engine(bind_ip="127.0.0.1", paths=2))
        m = MultiNodeCommunicator(nodes[k][l], nc)
        diag = ((k + 1) % N) * L + (l + 1) % L               # other box, other rail
        seen = []
        try:
            m.send(torch.zeros(4), diag)
        except NotImplementedError as e:
            seen.append("not routed" in str(e))
        try:
            m.all_reduce(torch.ones(4, dtype=torch.int32), "sum", scale=0.5)
        except ValueError:
            seen.append(True)
        try:
            nc.engine.isend(nc.flows[(k + 1) % N], torch.zeros(4, 4).t())
        except ValueError:
            seen.append(True)
        m.barrier()
        m.close()
        return seen

    for seen in _run_threads(W, fn):
        assert seen == [True, True, True]


def _async_ddp_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), UCCL_B200_NET_BIND_IP="127.0.0.1", UCCL_B200_NET_PATHS="2")
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from uccl_b200.parallel import AsyncMultiNode, MultiNodeCommunicator

    m = MultiNodeCommunicator.from_torch_dist(2, host=True, heap_bytes=128 << 20, stage_bytes=1 << 20, timeout_ms=60000)
    am = AsyncMultiNode(m)
    # several collectives in flight, completed in submission order
    xs = [torch.full((20_000,), float(rank + i)) for i in range(4)]
    ws = [am.all_reduce_async(x) for x in xs]
    ok = [all(bool((w.wait(60) == sum(r + i for r in range(world))).all()) for i, w in enumerate(ws))]
    # DDP with the asynchronous multi-box hook (small buckets -> several hook calls per backward)
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8))
    ddp = torch.nn.parallel.DistributedDataParallel(model, bucket_cap_mb=0.005)
    ddp.register_comm_hook(None, am.ddp_hook("avg"))
    torch.manual_seed(100 + rank)
    x = torch.randn(16, 64)
    ddp(x).pow(2).sum().backward()
    # reference: average of the per-rank gradients computed without DDP
    torch.manual_seed(0)
    ref = torch.nn.Sequential(torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 8))
    grads = []
    for r in range(world):
        ref.zero_grad()
        torch.manual_seed(100 + r)
        ref(torch.randn(16, 64)).pow(2).sum().backward()
        grads.append([p.grad.clone() for p in ref.parameters()])
    for i, p in enumerate(model.parameters()):
        ok.append(torch.allclose(p.grad, sum(g[i] for g in grads) / world, rtol=1e-4, atol=1e-5))
    am.close()
    m.barrier()
    q.put((rank, ok))
    dist.destroy_process_group()


def test_async_multibox_allreduce_and_ddp_hook():
    """Collectives queued on the helper thread complete in order, and DDP's bucketed gradient averaging across
    2 boxes x 2 ranks through the asynchronous hook equals the average of the single-rank gradients."""
    import multiprocessing as mp

    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_async_ddp_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in ps]
    for rank, ok in got:
        assert all(ok), (rank, ok)


def test_reordering_and_loss_together():
    """10 % of the datagrams are held back for 400 us (they arrive after hundreds of later packets, on every path)
    on top of 1 % loss: payloads stay intact, RACK's time-based window keeps spurious retransmissions modest."""
    a = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.01)
    b = net.Engine(bind_ip="127.0.0.1", paths=4, drop_prob=0.01)
    a.set_reorder(0.10, 400)
    b.set_reorder(0.10, 400)
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("fb", b.accept(lid)))
    t.start()
    fa = a.connect("127.0.0.1", b.port, lid)
    t.join()
    fb = box["fb"]
    g = torch.Generator().manual_seed(4)
    for n in (100, 70_000, 2_000_000):
        x = torch.randint(0, 255, (n,), generator=g, dtype=torch.uint8)
        y, z = torch.zeros_like(x), torch.zeros_like(x)
        w1, w2 = b.irecv(fb, y), a.irecv(fa, z)          # both directions at once
        s1, s2 = a.isend(fa, x), b.isend(fb, x)
        for w in (w1, w2, s1, s2):
            w.wait(120000)
        assert torch.equal(x, y) and torch.equal(x, z)
    st = a.flow_stats(fa)
    sb = b.flow_stats(fb)
    assert st["tx_pkts"] > 200
    # duplicates at the receiver = retransmissions that were not needed; they must stay a small fraction
    assert sb["rx_dup"] < 0.25 * sb["rx_pkts"], (sb["rx_dup"], sb["rx_pkts"])


def _deepep_mb_worker(rank, world, port, q):
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), UCCL_B200_LOCAL_SIZE="2", UCCL_B200_NET_BIND_IP="127.0.0.1",
                      UCCL_B200_NET_PATHS="2")
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import deep_ep

    buf = deep_ep.Buffer(dist.group.WORLD, num_nvl_bytes=1 << 20, num_rdma_bytes=1 << 20)   # how vLLM / SGLang build it
    T, H, K, E = 33, 128, 2, 8
    g = torch.Generator().manual_seed(50 + rank)
    x = torch.randn(T, H, generator=g).to(torch.bfloat16)
    idx = torch.rand(T, E, generator=g).topk(K, dim=1).indices.contiguous()
    w = torch.rand(T, K, generator=g)
    tpr, _, tpe, inr, _ = buf.get_dispatch_layout(idx, E)
    rx, ri, rw, pe, h, _ = buf.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=inr, num_tokens_per_expert=tpe, topk_idx=idx,
                                        topk_weights=w)
    comb, _, _ = buf.combine(rx, h, topk_weights=rw)
    ok = [type(buf).__name__ == "HostBuffer", buf.group_size == world, buf.get_num_rdma_ranks() == 2,
          torch.allclose(comb.float(), x.float() * inr.sum(1).float()[:, None], rtol=2e-2, atol=1e-1),
          sum(pe) == int(rx.size(0) and (ri >= 0).sum())]
    q.put((rank, ok))
    dist.destroy_process_group()


def test_deep_ep_buffer_from_a_process_group_that_spans_boxes():
    """`deep_ep.Buffer(group, ...)` exactly as the serving frameworks call it, with 4 ranks and 2 ranks per box."""
    import multiprocessing as mp

    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_deepep_mb_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    got = sorted(q.get(timeout=240) for _ in range(world))
    [p.join(60) for p in ps]
    for rank, ok in got:
        assert all(ok), (rank, ok)


@pytest.mark.parametrize("payload,inflight,cc", [(256, 8, "swift"), (60000, 240, "none"), (1400, 16, "eqds"), (1400, 248, "timely")])
def test_engine_configuration_corners(payload, inflight, cc):
    """Tiny and maximal datagrams, a window of 8 and of 248 packets, every congestion controller, unequal path counts,
    2 % loss: message sizes around the datagram boundary arrive intact in both directions at once."""
    a = net.Engine(bind_ip="127.0.0.1", paths=3, payload=payload, max_inflight=inflight, cc=cc, drop_prob=0.02)
    b = net.Engine(bind_ip="127.0.0.1", paths=5, payload=payload, max_inflight=inflight, cc=cc, drop_prob=0.02)
    lid = b.listen()
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("f", b.accept(lid)))
    t.start()
    fa = a.connect("127.0.0.1", b.port, lid)
    t.join()
    for n in (0, 1, payload - 1, payload, payload + 1, 70_001, 600_000 if payload > 1000 else 150_000):
        x = torch.randint(0, 255, (n,), dtype=torch.uint8)
        y, z = torch.zeros_like(x), torch.zeros_like(x)
        ws = [b.irecv(box["f"], y), a.irecv(fa, z), a.isend(fa, x), b.isend(box["f"], x)]
        for w in ws:
            w.wait(120000)
        assert torch.equal(x, y) and torch.equal(x, z), n
