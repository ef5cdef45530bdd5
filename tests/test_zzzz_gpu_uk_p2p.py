"""GPU test of the rank-addressed P2P communicator (`uccl_b200.ukernel.p2p`, the reference's ukernel_p2p surface --
experimental/ukernel/py/test_p2p.py, test_transport_paths.py): two ranks in one process on cuda:0, device tensors,
offset transfers in both directions at once.  Sorted last (written after the round's GPU budget)."""
import socket
import threading

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


def test_two_ranks_device_tensors():
    from uccl_b200.ukernel.p2p import Communicator

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    res, comms, errs = {}, {}, []
    n = 1 << 20

    def body(r):
        torch.cuda.set_device(0)
        c = Communicator(gpu_id=0, rank=r, world_size=2, exchanger_port=port)
        comms[r] = c
        peer = 1 - r
        assert c.accept_peer(peer) if r == 0 else c.connect_peer(peer)
        assert c.same_host(peer) and c.peer_transport(peer) == "ipc"
        recv = torch.zeros(n, dtype=torch.float32, device="cuda")
        send = torch.arange(n, dtype=torch.float32, device="cuda") + 7 * (r + 1)
        torch.cuda.synchronize()
        assert c.reg_rdma(10 + r, recv) and c.wait_mr(peer, 10 + peer)
        rq = c.irecv(peer, recv, offset=4096, len=(n // 2) * 4)
        sq = c.isend(peer, send, offset=1024, len=(n // 2) * 4, remote_buffer_id=10 + peer, remote_offset=4096)
        assert c.wait_finish_multi([rq, sq])
        torch.cuda.synchronize()
        res[r] = recv.cpu()
        assert c.barrier()
        assert c.unreg_rdma(10 + r)

    def run(r):
        try:
            body(r)
        except Exception as e:  # pragma: no cover
            import traceback

            traceback.print_exc()
            errs.append(e)

    ths = [threading.Thread(target=run, args=(r,), daemon=True) for r in range(2)]
    [t.start() for t in ths]
    [t.join(120) for t in ths]
    assert not errs, errs
    assert not any(t.is_alive() for t in ths), "a rank is stuck"
    for r in range(2):
        exp = torch.zeros(n)
        exp[1024:1024 + n // 2] = torch.arange(256, 256 + n // 2, dtype=torch.float32) + 7 * (2 - r)
        assert torch.equal(res[r], exp)
    for r in (1, 0):
        comms[r].close()
