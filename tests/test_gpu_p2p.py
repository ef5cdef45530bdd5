"""GPU tests of the P2P transfer engine: same-process endpoints (direct pointers) and two real
processes on one GPU (CUDA-IPC mapped peer memory) -- fill-pattern oracles in the style of the
reference's p2p/tests/test_engine_{send,read,write}.py and test_engine_onesided_ipc.py."""
import multiprocessing as mp

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def _pair():
    from uccl_b200.p2p import Endpoint

    ng = torch.cuda.device_count()
    a = Endpoint(0)
    b = Endpoint(1 if ng > 1 else 0)
    ok, conn = a.connect(remote_metadata=b.get_metadata())
    assert ok
    ok2, ip, gpu, conn_b = b.accept(5000)
    assert ok2 and gpu == 0
    return a, b, conn, conn_b


def test_metadata_roundtrip():
    from uccl_b200.p2p import Endpoint

    e = Endpoint(0)
    ip, port, gpu = Endpoint.parse_metadata(e.get_metadata())
    assert ip == "127.0.0.1" and port > 0 and gpu == 0


@pytest.mark.parametrize("nbytes", [1, 100, 4096, 1 << 20, (8 << 20) + 13])
def test_send_recv_same_process(nbytes):
    a, b, conn, conn_b = _pair()
    src = torch.randint(0, 255, (nbytes,), dtype=torch.uint8, device=f"cuda:{a.local_gpu_idx}")
    dst = torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{b.local_gpu_idx}")
    torch.cuda.synchronize()
    ok, rt = b.recv_async(conn_b, 0, dst.data_ptr(), nbytes)
    assert ok
    assert a.send(conn, 0, src.data_ptr(), nbytes)
    assert b.wait(rt, 10000)
    assert torch.equal(dst.cpu(), src.cpu())


@pytest.mark.parametrize("strategy", ["none", "for"])
def test_send_recv_compressed(strategy):
    """Compression hook: header + (optionally FoR-compressed) payload; the receiver always ends up
    with the sender's exact bits."""
    import threading

    a, b, conn, conn_b = _pair()
    src = (torch.randn(3 << 20, device=f"cuda:{a.local_gpu_idx}") * 2).to(torch.bfloat16)  # 6 MiB > threshold
    dst = torch.zeros_like(src, device=f"cuda:{b.local_gpu_idx}")
    torch.cuda.synchronize()
    res = {}

    def receiver():
        torch.cuda.set_device(b.local_gpu_idx)
        res["ok"] = b.recv_compressed(conn_b, dst)

    t = threading.Thread(target=receiver)
    t.start()
    before = a.stats()["bytes_sent"]
    assert a.send_compressed(conn, src, strategy=strategy)
    t.join(60)
    assert res.get("ok")
    torch.cuda.synchronize()
    assert torch.equal(dst.view(torch.int16).cpu(), src.view(torch.int16).cpu())
    sent = a.stats()["bytes_sent"] - before
    if strategy == "for":
        assert sent < 0.9 * src.numel() * 2, sent  # the wire really carried fewer bytes
    else:
        assert sent == src.numel() * 2 + 16


def test_onesided_vector_write_read():
    a, b, conn, conn_b = _pair()
    dev_a, dev_b = f"cuda:{a.local_gpu_idx}", f"cuda:{b.local_gpu_idx}"
    # "KV blocks": distinct value per iov, odd sizes to exercise the tail path
    sizes = [128 << 10, (256 << 10) + 7, 4096, 33, 1 << 20]
    srcs = [torch.full((s,), i + 1, dtype=torch.uint8, device=dev_a) for i, s in enumerate(sizes)]
    dsts = [torch.zeros(s, dtype=torch.uint8, device=dev_b) for s in sizes]
    descs_b = b.register_memory(dsts)
    blob = b.get_serialized_descs(descs_b)
    remote = a.deserialize_descs(blob)
    local = a.register_memory(srcs)
    torch.cuda.synchronize()
    ok, tid = a.transfer(conn, "write", local, remote)
    assert ok and a.wait(tid, 10000)
    for i, d in enumerate(dsts):
        assert bool((d == i + 1).all())
    # read them back into fresh buffers
    back = [torch.zeros(s, dtype=torch.uint8, device=dev_a) for s in sizes]
    lb = a.register_memory(back)
    ok, tid = a.transfer(conn, "read", lb, remote)
    assert ok
    done = False
    for _ in range(2000000):
        ok, done = a.poll_async(tid)
        assert ok
        if done:
            break
    assert done
    for i, d in enumerate(back):
        assert bool((d == i + 1).all())
    st = a.stats()
    assert st["kernel_launches"] >= 2 and st["memcpy_fallbacks"] == 0


def test_notifications():
    a, b, conn, conn_b = _pair()
    assert a.send_notif(conn, b"kv-ready:42")
    import time

    got = []
    for _ in range(200):
        got = b.get_notifs()
        if got:
            break
        time.sleep(0.01)
    assert got and got[0][1] == b"kv-ready:42"


def _server(q_md, q_res, nbytes):
    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(0)
    e = Endpoint(0)
    q_md.put(e.get_metadata())
    ok, ip, gpu, conn = e.accept(60000)
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda:0")
    win = torch.full((nbytes,), 7, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    # two-sided receive, then advertise a window for the client's one-sided read
    okr = e.recv(conn, 0, buf.data_ptr(), nbytes)
    desc = e.register_memory([win])
    e.send_notif(conn, e.get_serialized_descs(desc))
    # wait for the client's "done" notification
    import time

    t0 = time.time()
    fin = False
    while time.time() - t0 < 60 and not fin:
        for _, m in e.get_notifs():
            if m == b"done":
                fin = True
        time.sleep(0.005)
    q_res.put((bool(ok), bool(okr), int(buf.sum().item()), fin))


def _client(q_md, q_res, nbytes):
    import time

    import torch

    from uccl_b200.p2p import Endpoint

    torch.cuda.set_device(0)
    e = Endpoint(0)
    md = q_md.get(timeout=60)
    ok, conn = e.connect(remote_metadata=md)
    src = torch.ones(nbytes, dtype=torch.uint8, device="cuda:0")
    torch.cuda.synchronize()
    oks = e.send(conn, 0, src.data_ptr(), nbytes)
    blob = None
    t0 = time.time()
    while blob is None and time.time() - t0 < 60:
        for _, m in e.get_notifs():
            blob = m
        time.sleep(0.005)
    remote = e.deserialize_descs(blob)
    dst = torch.zeros(nbytes, dtype=torch.uint8, device="cuda:0")
    okr = e.read(conn, 0, dst.data_ptr(), nbytes, remote[0])
    val = int(dst.sum().item())
    e.send_notif(conn, b"done")
    q_res.put((bool(ok), bool(oks), bool(okr), val))


def test_two_processes_cuda_ipc():
    nbytes = (2 << 20) + 48
    ctx = mp.get_context("spawn")
    q_md, q_s, q_c = ctx.Queue(), ctx.Queue(), ctx.Queue()
    ps = [ctx.Process(target=_server, args=(q_md, q_s, nbytes)), ctx.Process(target=_client, args=(q_md, q_c, nbytes))]
    [p.start() for p in ps]
    rs = q_s.get(timeout=180)
    rc = q_c.get(timeout=180)
    [p.join(30) for p in ps]
    assert rs == (True, True, nbytes, True)
    assert rc == (True, True, True, 7 * nbytes)


def test_prepared_transfer_moves_thousands_of_blocks_in_one_launch():
    """prepare_transfer / post_transfer (NIXL prepXfer / postXfer): 2000 KV blocks -- more than fit in the kernel
    parameters, so they travel in a pinned descriptor table -- move with ONE launch per post, in both directions,
    repeatedly, incl. blocks whose size is not a multiple of 16 bytes."""
    import torch

    from uccl_b200.p2p import Endpoint

    d0, d1 = 0, (1 if torch.cuda.device_count() > 1 else 0)
    a, b = Endpoint(d0), Endpoint(d1)
    ok, conn = a.connect(remote_metadata=b.get_metadata())
    assert ok
    b.accept(5000)
    nb, blk = 2000, 4096 + 8
    g = torch.Generator().manual_seed(1)
    src = torch.randint(0, 255, (nb, blk), dtype=torch.uint8, generator=g).to(f"cuda:{d0}")
    dst = torch.zeros(nb, blk, dtype=torch.uint8, device=f"cuda:{d1}")
    back = torch.zeros(nb, blk, dtype=torch.uint8, device=f"cuda:{d0}")
    perm = torch.randperm(nb, generator=g).tolist()  # scatter: block i lands in row perm[i]
    la = a.register_memory([src[i] for i in range(nb)])
    ra = a.deserialize_descs(b.get_serialized_descs(b.register_memory([dst[perm[i]] for i in range(nb)])))
    lback = a.register_memory([back[i] for i in range(nb)])
    torch.cuda.synchronize(d0)
    torch.cuda.synchronize(d1)
    k0 = a.stats()["kernel_launches"]
    wr = a.prepare_transfer(conn, "write", la, ra)
    rd = a.prepare_transfer(conn, "read", lback, ra)
    for it in range(3):
        ok, tid = a.post_transfer(wr)
        assert ok and a.wait(tid)
        ok, tid = a.post_transfer(rd)
        assert ok and a.wait(tid)
    assert a.stats()["kernel_launches"] - k0 == 6  # one launch per post
    torch.cuda.synchronize(d0)
    torch.cuda.synchronize(d1)
    assert torch.equal(dst[perm].cpu(), src.cpu())
    assert torch.equal(back.cpu(), src.cpu())
    assert a.release_transfer(wr) and a.release_transfer(rd) and not a.release_transfer(wr)
