"""Python-level grouped send/recv on the GPU (`Communicator.batch_send_recv`): same kernel and the same
patterns as the C++ NCCL-API test (ring step with a multi-chunk message, all-to-all incl. self), through
the pybind path; plus the EP Buffer's optional arguments and the MoE training step on the CUDA kernels.

The same assertions also run against the CPU reference backends (tests/test_host_ep.py,
test_host_collectives.py)."""
import pytest
import torch

from helpers import get_world, run_ranks

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(90)]


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
    # same expert choice as the layer (its router runs in bf16: a near-tie could pick another expert in fp32)
    with torch.no_grad():
        idx = torch.topk(F.softmax(m.router(x.detach()).float(), dim=-1), K, dim=-1).indices
    w = F.softmax(xr @ rw.T, dim=-1).gather(1, idx)
    yr = torch.zeros(T, H, device=dev)
    for e in range(E):
        sel = idx == e
        rows = sel.any(1).nonzero().flatten()
        if rows.numel():
            yr = yr.index_add(0, rows, (F.silu(xr[rows] @ w1[e]) @ w2[e]) * (w * sel).sum(1)[rows][:, None])
    (yr * g.float()).sum().backward()

    def close(a, b, tol):
        # the layer runs bf16 GEMMs / bf16 accumulation of expert outputs, the oracle is fp32: compare in the
        # Frobenius norm (a single bf16-rounded element may be off by more than a max-abs bound allows)
        a, b = a.float(), b.float()
        return ((a - b).norm() / b.norm().clamp_min(1e-6)).item() <= tol

    assert close(y, yr.detach(), 5e-2)
    assert close(x.grad, xr.grad, 8e-2)
    assert close(m.w1.grad, w1.grad, 8e-2) and close(m.w2.grad, w2.grad, 8e-2)
    assert close(m.router.weight.grad, rw.grad, 1e-1)


@pytest.mark.parametrize("n", [2, 4])
def test_ep_buffer_optional_arguments(n):
    """expert_alignment, num_worst_tokens, combine bias (one and two tensors) and the
    async_finish / previous_event stream choreography of the normal-mode kernels -- the same expectations as
    tests/test_host_ep.py::test_host_ep_dispatch_combine has for the CPU backend."""
    from test_gpu_ep import get_buffers, make_inputs, ref_layout, run_threads

    T, H, K = 97, 512, 3
    E = n * 2
    e_per = E // n
    bufs = get_buffers(n)
    xs, idxs, ws = make_inputs(n, T, H, K, E, seed=40 + n)
    layouts = [ref_layout(idxs[r], n, E) for r in range(n)]
    g = torch.Generator().manual_seed(5)
    b0s = [torch.randn(T, H, generator=g).to(torch.bfloat16) for _ in range(n)]
    b1s = [torch.randn(T, H, generator=g).to(torch.bfloat16) for _ in range(n)]

    def fn(b):
        r, dev = b.rank, b.device
        x, idx, w = xs[r].to(dev), idxs[r].to(dev), ws[r].to(dev)
        b0, b1 = b0s[r].to(dev), b1s[r].to(dev)
        tpr, _, tpe, in_rank, ev0 = b.get_dispatch_layout(idx, E, async_finish=True)
        assert ev0.event is not None
        rx, ri, rw, pe, h, ev1 = b.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank,
                                            num_tokens_per_expert=tpe, topk_idx=idx, topk_weights=w,
                                            expert_alignment=8, previous_event=ev0, async_finish=True)
        ev1.current_stream_wait()
        num_recv = rx.size(0)
        rx_keep = rx.clone()
        cb = b.get_combine_buffer(num_recv, H, K)
        cb.copy_(rx)
        with b.combine(cb, h, topk_weights=rw, bias=(b0, b1), async_finish=True)[2]:
            pass  # EventOverlap as a context manager: the current stream waits on exit
        comb2, cw2, ev2 = b.combine(cb, h, topk_weights=rw, bias=(b0, b1), async_finish=True)
        ev2.current_stream_wait()
        comb1, _, _ = b.combine(cb, h, bias=b0)
        torch.cuda.current_stream().synchronize()
        comb2, cw2, comb1 = comb2.cpu(), cw2.cpu(), comb1.cpu()
        # CUDA-graph friendly variant: fixed-size outputs, no host wait, -1 padded expert ids
        wx, wi, ww, wpe, wh, _ = b.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=in_rank,
                                            num_tokens_per_expert=tpe, topk_idx=idx, topk_weights=w,
                                            num_worst_tokens=n * T)
        torch.cuda.current_stream().synchronize()
        return dict(rx=rx_keep.cpu(), pe=pe, comb2=comb2, cw2=cw2, comb1=comb1, wx=wx.cpu(), wi=wi.cpu(), wpe=wpe,
                    num_recv=num_recv)

    outs = run_threads(bufs, fn)
    for r, o in enumerate(outs):
        exp_rows = torch.cat([xs[s][layouts[s][2][:, r].nonzero().flatten()] for s in range(n)])
        assert torch.equal(o["rx"], exp_rows)
        counts = [int(sum((idxs[s] == r * e_per + e).sum() for s in range(n))) for e in range(e_per)]
        assert o["pe"] == [(c + 7) // 8 * 8 for c in counts]
        fan = layouts[r][2].sum(1).float()[:, None]
        base = xs[r].float() * fan
        assert torch.allclose(o["comb1"].float(), base + b0s[r].float(), rtol=2e-2, atol=2e-1)
        assert torch.allclose(o["comb2"].float(), base + b0s[r].float() + b1s[r].float(), rtol=2e-2, atol=2e-1)
        assert torch.allclose(o["cw2"], torch.where(idxs[r] >= 0, ws[r], torch.zeros_like(ws[r])), rtol=1e-5, atol=1e-6)
        assert o["wx"].size(0) == n * T and o["wpe"] == []
        assert torch.equal(o["wx"][:o["num_recv"]], exp_rows)
        assert bool((o["wi"][o["num_recv"]:] == -1).all())


_NCCL_PLUGIN_SCRIPT = r"""
import os, torch, torch.distributed as dist
dist.init_process_group("nccl")
rank = dist.get_rank()
torch.cuda.set_device(rank)
x = torch.full((1 << 20,), float(rank + 1), device="cuda")
dist.all_reduce(x)
torch.cuda.synchronize()
w = dist.get_world_size()
assert bool((x == w * (w + 1) / 2).all())
y = [torch.empty(1000, device="cuda") for _ in range(w)]
dist.all_gather(y, torch.full((1000,), float(rank), device="cuda"))
assert all(bool((y[r] == r).all()) for r in range(w))
print(f"rank {rank} ok", flush=True)
dist.destroy_process_group()
"""


def test_nccl_runs_over_the_net_plugin(tmp_path):
    """Stock NCCL (torch's) with NVLink P2P and SHM disabled is forced onto its network transport, and
    NCCL_NET_PLUGIN points it at libnccl-net-uccl_b200.so: the all-reduce then crosses our multipath
    datagram transport (host staged).  Needs 2 GPUs; this is the single-box stand-in for two boxes."""
    import os
    import subprocess
    import sys

    from uccl_b200 import _build

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _build.build()
    plugin = _build.nccl_net_plugin_path()
    script = tmp_path / "nccl_net.py"
    script.write_text(_NCCL_PLUGIN_SCRIPT)
    env = dict(os.environ, NCCL_NET_PLUGIN=str(plugin), NCCL_P2P_DISABLE="1", NCCL_SHM_DISABLE="1", NCCL_DEBUG="INFO",
               NCCL_DEBUG_SUBSYS="INIT,NET", UCCL_B200_NET_IFNAME="lo", NCCL_IB_DISABLE="1",
               LD_LIBRARY_PATH=str(plugin.parent) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), str(script)],
                       capture_output=True, text=True, timeout=80, env=env)
    out = r.stdout + r.stderr
    sys.stdout.write(out[-3000:])
    assert r.returncode == 0
    assert out.count(" ok") >= 2
    assert "uccl_b200" in out  # NCCL logged the plugin it selected


def test_p2p_net_channel_gpu_tensors():
    """Cross-box P2P path with CUDA tensors on both ends (pinned staging pipelines), over loopback."""
    import threading

    from uccl_b200 import net
    from uccl_b200.p2p import NetChannel

    dev = torch.device("cuda", 0)
    ea, eb = net.Engine(bind_ip="127.0.0.1", paths=4), net.Engine(bind_ip="127.0.0.1", paths=4)
    srv = NetChannel.listen(eb, chunk_bytes=1 << 20)
    box = {}
    t = threading.Thread(target=lambda: box.setdefault("ch", srv.accept()))
    t.start()
    cli = NetChannel.connect(ea, ("127.0.0.1", eb.port, srv.address[2]), chunk_bytes=1 << 20)
    t.join()
    blocks = [torch.randn(3_000_001, device=dev), torch.randn(1000, device=dev).to(torch.bfloat16)]
    outs = [torch.zeros_like(b) for b in blocks]

    def rx():
        torch.cuda.set_device(0)
        box["n"] = srv.recv_tensors(outs)

    rt = threading.Thread(target=rx)
    rt.start()
    sent = cli.send_tensors(blocks)
    rt.join()
    torch.cuda.synchronize()
    assert box["n"] == sent
    for a, b in zip(blocks, outs):
        assert torch.equal(a, b)


def test_ep_buffer_across_boxes_with_gpu_tensors():
    """Two 'boxes' of one GPU rank each (same physical GPU), datagram rail between them: the portable
    multi-box EP path accepts CUDA tensors and returns CUDA results."""
    import threading

    from uccl_b200 import Communicator, net
    from uccl_b200.ep import Buffer
    from uccl_b200.parallel import MultiNodeCommunicator

    dev = torch.device("cuda", 0)
    W, T, H, K, E = 2, 64, 256, 2, 4
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(T, H, generator=g).to(torch.bfloat16) for _ in range(W)]
    idxs = [torch.rand(T, E, generator=g).topk(K, dim=1).indices.contiguous() for _ in range(W)]
    ws = [torch.rand(T, K, generator=g) for _ in range(W)]
    locals_ = [Communicator.local_world(1, devices=[0], heap_bytes=256 << 20, stage_bytes=8 << 20, timeout_ms=5000)[0] for _ in range(W)]
    slots, bar = [None] * W, threading.Barrier(W)

    def exchange_for(r):
        def ex(obj):
            slots[r] = obj
            bar.wait()
            out = list(slots)
            bar.wait()
            return out
        return ex

    outs, errs = [None] * W, []

    def fn(r):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(device=dev)):
                m = MultiNodeCommunicator(locals_[r], net.NetCommunicator(r, W, exchange_for(r), engine=net.Engine(bind_ip="127.0.0.1", paths=2)))
                b = Buffer(comm=m, num_nvl_bytes=1 << 20)
                x, idx, w = xs[r].to(dev), idxs[r].to(dev), ws[r].to(dev)
                tpr, _, tpe, inr, _ = b.get_dispatch_layout(idx, E)
                rx, ri, rw, pe, h, _ = b.dispatch(x, num_tokens_per_rank=tpr, is_token_in_rank=inr, num_tokens_per_expert=tpe,
                                                  topk_idx=idx, topk_weights=w)
                comb, _, _ = b.combine(rx, h, topk_weights=rw)
                assert rx.is_cuda and comb.is_cuda
                outs[r] = (rx.cpu(), inr.cpu(), comb.cpu())
                m.close()
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=fn, args=(r,)) for r in range(W)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    for r, (rx, inr, comb) in enumerate(outs):
        exp = torch.cat([xs[s][outs[s][1][:, r].nonzero().flatten()] for s in range(W)])
        assert torch.equal(rx, exp)
        assert torch.allclose(comb.float(), xs[r].float() * inr.sum(1).float()[:, None], rtol=2e-2, atol=1e-1)


def test_nccl_api_across_boxes_with_device_buffers(tmp_path):
    """The NCCL drop-in with more ranks than the configured box size (2 processes x 1 GPU rank on device 0):
    MultiComm's pinned staging + rail ring with cudaMalloc'd buffers."""
    import os
    import subprocess
    import sys

    from uccl_b200 import _build

    _build.build()
    shim = _build.nccl_shim_path()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "nccl_multibox_gpu_test"
    subprocess.run(["g++", "-std=c++17", "-O1", os.path.join(root, "tests/cpp/nccl_multibox_gpu_test.cc"), "-I/usr/include",
                    "-I/usr/local/cuda/include", "-L" + str(shim.parent), "-luccl_b200_nccl", "-Wl,-rpath," + str(shim.parent),
                    "-L/usr/local/cuda/lib64", "-lcudart", "-lpthread", "-o", str(exe)], check=True)
    r = subprocess.run([str(exe), "2"], capture_output=True, text=True, timeout=80)
    sys.stdout.write(r.stdout + r.stderr[-2000:])
    assert r.returncode == 0 and "nccl_multibox_gpu_test: OK" in r.stdout


def test_proxy_forwards_device_commands_to_another_box():
    """GPU kernel -> D2H command queue -> CPU proxy -> datagram transport -> remote proxy -> remote GPU heap +
    signal: two 'boxes' of one GPU rank each (same device), destinations addressed by global rank."""
    import threading
    import time

    from uccl_b200 import Communicator, net
    from uccl_b200.ep.proxy import Proxy

    W = 2
    comms = [Communicator.local_world(1, devices=[0], heap_bytes=256 << 20, stage_bytes=8 << 20, timeout_ms=5000)[0] for _ in range(W)]
    slots, bar = [None] * W, threading.Barrier(W)

    def exchange_for(r):
        def ex(obj):
            slots[r] = obj
            bar.wait()
            out = list(slots)
            bar.wait()
            return out
        return ex

    res, errs = [None] * W, []

    def fn(b):
        try:
            torch.cuda.set_device(0)
            with torch.cuda.stream(torch.cuda.Stream(device=0)):
                c = comms[b]
                rail = net.NetCommunicator(b, W, exchange_for(b), engine=net.Engine(bind_ip="127.0.0.1", paths=2))
                inbox = c.zeros(1 << 16, dtype=torch.uint8)
                counter = c.zeros(1, dtype=torch.int64)
                src = c.empty(1 << 16, dtype=torch.uint8)
                src.fill_(b + 7)
                torch.cuda.current_stream().synchronize()
                p = Proxy(c, rail=rail)
                bar.wait()
                other = 1 - b
                p.device_write(other, src, c.native.heap_offset(inbox.data_ptr()),
                               signal_offset=c.native.heap_offset(counter.data_ptr()), signal_value=5)
                t0 = time.time()
                while int(counter.item()) != 5:
                    assert time.time() - t0 < 20
                    time.sleep(0.001)
                res[b] = bool((inbox == other + 7).all())
                bar.wait()
                p.stop()
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=fn, args=(b,)) for b in range(W)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    assert not errs, errs
    assert res == [True, True]


def _free_port() -> int:
    import socket

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _xhost_server(q_md, q_res, nbytes):
    import os
    import time

    os.environ["UCCL_B200_P2P_HOST_ID"] = "1001"    # pretend to be another machine than the client
    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(0)
    e = Endpoint(0)
    q_md.put(e.get_metadata())
    ok, ip, gpu, conn = e.accept(60000)
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    win = torch.full((nbytes,), 7, dtype=torch.uint8, device="cuda")
    inbox = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    okr = e.recv(conn, 0, buf.data_ptr(), nbytes)
    descs = e.register_memory([win, inbox])
    e.send_notif(conn, e.get_serialized_descs(descs))
    t0, fin = time.time(), False
    while time.time() - t0 < 60 and not fin:
        for _, m in e.get_notifs():
            fin = fin or m == b"done"
        time.sleep(0.002)
    torch.cuda.synchronize()
    q_res.put((bool(ok), bool(okr), int(buf.sum().item()), int(inbox.to(torch.int64).sum().item()), fin))


def _xhost_client(q_md, q_res, nbytes):
    import os
    import time

    os.environ["UCCL_B200_P2P_HOST_ID"] = "1002"
    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(0)
    e = Endpoint(0)
    ok, conn = e.connect(remote_metadata=q_md.get(timeout=60))
    src = torch.ones(nbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    oks = e.send(conn, 0, src.data_ptr(), nbytes)
    blob, t0 = None, time.time()
    while blob is None and time.time() - t0 < 60:
        for _, m in e.get_notifs():
            blob = m
        time.sleep(0.002)
    remote = e.deserialize_descs(blob)
    dst = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    okr = e.read(conn, 0, dst.data_ptr(), nbytes, remote[0])
    three = torch.full((nbytes,), 3, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    okw = e.write(conn, 0, three.data_ptr(), nbytes, remote[1])
    e.send_notif(conn, b"done")
    torch.cuda.synchronize()
    q_res.put((bool(ok), bool(oks), bool(okr), int(dst.to(torch.int64).sum().item()), bool(okw)))
    time.sleep(0.3)


def test_p2p_endpoint_between_hosts_stages_gpu_memory():
    """Two GPU endpoints that believe they sit on different machines (UCCL_B200_P2P_HOST_ID): send/recv, one-sided
    read and write of device memory travel over the connection, staged through host bounce buffers on both ends."""
    import multiprocessing as mp

    nbytes = (5 << 20) + 3
    ctx = mp.get_context("spawn")
    q_md, q_s, q_c = ctx.Queue(), ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_xhost_server, args=(q_md, q_s, nbytes)), ctx.Process(target=_xhost_client, args=(q_md, q_c, nbytes))]
    [p.start() for p in ps]
    rs = q_s.get(timeout=80)
    rc = q_c.get(timeout=80)
    [p.join(30) for p in ps]
    assert rs == (True, True, nbytes, 3 * nbytes, True)
    assert rc == (True, True, True, 7 * nbytes, True)
