"""Rank-addressed point-to-point communicator: the reference's ``ukernel_p2p.Communicator`` surface
(experimental/ukernel/py/ukernel_p2p.cpp:407-444 -- connect / accept by rank, buffer ids published through an
exchanger, isend / irecv with byte offsets, request polling, named barriers).

Here it is a thin layer over this library's P2P engine (``uccl_b200.p2p.Endpoint``: copy kernels over CUDA IPC inside a
box, the TCP data path otherwise, host memory in GPU-less CI) plus a small exchanger -- a key/value rendezvous served
by rank 0 that carries endpoint metadata, published buffer ids and barrier arrivals (the reference's socket OOB
exchanger, experimental/ukernel/src/transport/oob).

    comm = Communicator(gpu_id=0, rank=r, world_size=2, exchanger_port=29610)
    comm.accept_peer(1) if r == 0 else comm.connect_peer(0)
    req = comm.isend(peer, t, offset=256, len=1024); comm.wait_finish(req)
"""
from __future__ import annotations

import socket
import threading
import time
from typing import Any, Dict, List, Optional

import torch

from ..p2p import Endpoint
from ..p2p.utils import create_socket_and_connect, recv_obj, send_obj

__all__ = ["Communicator", "Exchanger"]

_TRANSPORTS = ("auto", "ipc", "uccl", "tcp")


class Exchanger:
    """Key/value rendezvous.  ``put`` stores, ``get`` blocks until the key exists (or the timeout passes),
    ``count`` blocks until `n` keys with a prefix exist.  One thread per client connection; values are pickles, so --
    like torch.distributed's object collectives -- bind it to an address only trusted peers can reach."""

    def __init__(self, ip: str, port: int):
        self._kv: Dict[str, Any] = {}
        self._cv = threading.Condition()
        self._srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self._srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self._srv.bind((ip, int(port)))
        self._srv.listen(64)
        self.port = self._srv.getsockname()[1]
        self._stop = False
        self._threads: List[threading.Thread] = []
        self._acceptor = threading.Thread(target=self._accept_loop, name="uk-exchanger", daemon=True)
        self._acceptor.start()

    def _accept_loop(self):
        while not self._stop:
            try:
                c, _ = self._srv.accept()
            except OSError:
                return
            t = threading.Thread(target=self._serve, args=(c,), daemon=True)
            t.start()
            self._threads.append(t)

    def _wait(self, pred, timeout_ms):
        deadline = None if timeout_ms is None or timeout_ms < 0 else time.monotonic() + timeout_ms / 1e3
        with self._cv:
            while not pred():
                left = None if deadline is None else deadline - time.monotonic()
                if self._stop or (left is not None and left <= 0):
                    return False
                self._cv.wait(0.2 if left is None else min(0.2, left))
            return True

    def _serve(self, c: socket.socket):
        try:
            while True:
                msg = recv_obj(c)
                op = msg[0]
                if op == "put":
                    with self._cv:
                        self._kv[msg[1]] = msg[2]
                        self._cv.notify_all()
                    send_obj(c, True)
                elif op == "get":
                    ok = self._wait(lambda: msg[1] in self._kv, msg[2])
                    send_obj(c, (ok, self._kv.get(msg[1])))
                elif op == "count":
                    ok = self._wait(lambda: sum(1 for k in self._kv if k.startswith(msg[1])) >= msg[2], msg[3])
                    send_obj(c, ok)
                elif op == "del":
                    with self._cv:
                        self._kv.pop(msg[1], None)
                    send_obj(c, True)
                else:
                    send_obj(c, False)
        except (ConnectionError, OSError, EOFError):
            pass
        finally:
            c.close()

    def close(self):
        self._stop = True
        with self._cv:
            self._cv.notify_all()
        try:
            self._srv.close()
        except OSError:
            pass


class _ExchangerClient:
    def __init__(self, ip: str, port: int, timeout_s: float = 60.0):
        retries = max(1, int(timeout_s / 0.2))
        self._s = create_socket_and_connect(ip, port, max_retries=retries, initial_delay=0.05, backoff=1.5, max_delay=0.2)
        self._mu = threading.Lock()

    def _call(self, *msg):
        with self._mu:
            send_obj(self._s, msg)
            return recv_obj(self._s)

    def put(self, key: str, value: Any) -> None:
        self._call("put", key, value)

    def get(self, key: str, timeout_ms: int = -1):
        return self._call("get", key, timeout_ms)

    def count(self, prefix: str, n: int, timeout_ms: int = -1) -> bool:
        return bool(self._call("count", prefix, n, timeout_ms))

    def delete(self, key: str) -> None:
        self._call("del", key)

    def close(self):
        try:
            self._s.close()
        except OSError:
            pass


class Communicator:
    """See the module docstring.  Methods return ``bool`` / request ids like the reference's binding; argument errors
    raise ``ValueError``."""

    def __init__(self, gpu_id: int, rank: int, world_size: int, exchanger_ip: str = "127.0.0.1",
                 exchanger_port: int = 6979, transport: str = "auto", local_id: int = -1):
        if transport not in _TRANSPORTS:
            raise ValueError(f"transport must be one of {_TRANSPORTS}, got {transport!r}")
        if not 0 <= int(rank) < int(world_size):
            raise ValueError(f"rank {rank} outside a world of {world_size}")
        self._rank, self._world = int(rank), int(world_size)
        self._transport = transport
        self._server = Exchanger(exchanger_ip, exchanger_port) if self._rank == 0 else None
        self._x = _ExchangerClient(exchanger_ip, exchanger_port)
        use_gpu = torch.cuda.is_available() and int(gpu_id) >= 0
        if use_gpu:
            torch.cuda.set_device(int(gpu_id))
        self._ep = Endpoint(int(gpu_id) if use_gpu else -1)
        self._gpu = int(gpu_id) if use_gpu else -1
        self._host = socket.gethostname()
        self._x.put(f"ep/{self._rank}", {"md": self._ep.get_metadata(), "host": self._host, "gpu": self._gpu,
                                        "local_id": int(local_id)})
        self._conns: Dict[int, int] = {}
        self._peers: Dict[int, dict] = {}
        self._bind: Dict[int, tuple] = {}      # data_ptr -> (tensor, buffer_id, bytes, mr, kind)
        self._ids: Dict[int, int] = {}         # buffer_id -> data_ptr
        self._pending: Dict[int, tuple] = {}   # req -> (transfer id, tensor, temporary mr or None)
        self._next_req = 1
        self._barrier_seq: Dict[str, int] = {}
        self._mu = threading.Lock()
        self._closed = False

    # ------------------------------------------------------------------ identity
    @property
    def rank(self) -> int:
        return self._rank

    @property
    def world_size(self) -> int:
        return self._world

    def _peer(self, peer_rank: int) -> dict:
        p = self._peers.get(peer_rank)
        if p is None:
            ok, p = self._x.get(f"ep/{peer_rank}", 60000)
            if not ok:
                raise RuntimeError(f"rank {peer_rank} never published its endpoint")
            self._peers[peer_rank] = p
        return p

    def same_host(self, peer_rank: int) -> bool:
        return self._peer(int(peer_rank))["host"] == self._host

    def peer_transport(self, peer_rank: int) -> str:
        """"ipc" for a peer GPU in this box (copy kernels over mapped memory), otherwise "tcp"."""
        p = self._peer(int(peer_rank))
        if self._transport in ("ipc", "tcp"):
            return self._transport
        return "ipc" if (p["host"] == self._host and self._gpu >= 0 and p["gpu"] >= 0) else "tcp"

    # ------------------------------------------------------------------ connections
    def connect_peer(self, peer_rank: int) -> bool:
        peer_rank = int(peer_rank)
        if peer_rank == self._rank or not 0 <= peer_rank < self._world:
            raise ValueError(f"bad peer rank {peer_rank}")
        if peer_rank in self._conns:
            return True
        p = self._peer(peer_rank)
        # the acceptor serves one peer at a time and says whom it is waiting for: no cross-matched accepts
        ok, _ = self._x.get(f"accepting/{peer_rank}/{self._rank}", 120000)
        if not ok:
            return False
        ok, conn = self._ep.connect(remote_metadata=p["md"])
        if not ok:
            return False
        self._conns[peer_rank] = conn
        self._ep.set_rank_conn(peer_rank, conn)
        self._x.put(f"connected/{self._rank}/{peer_rank}", True)
        return True

    def accept_peer(self, peer_rank: int) -> bool:
        peer_rank = int(peer_rank)
        if peer_rank == self._rank or not 0 <= peer_rank < self._world:
            raise ValueError(f"bad peer rank {peer_rank}")
        if peer_rank in self._conns:
            return True
        self._x.put(f"accepting/{self._rank}/{peer_rank}", True)
        ok, _ip, _gpu, conn = self._ep.accept(120000)
        if not ok:
            return False
        self._conns[peer_rank] = conn
        self._ep.set_rank_conn(peer_rank, conn)
        self._x.get(f"connected/{peer_rank}/{self._rank}", 120000)
        self._x.delete(f"accepting/{self._rank}/{peer_rank}")
        return True

    def _conn(self, peer_rank: int) -> int:
        c = self._conns.get(int(peer_rank))
        if c is None:
            raise RuntimeError(f"no connection to rank {peer_rank}: call connect_peer / accept_peer first")
        return c

    # ------------------------------------------------------------------ registration
    @staticmethod
    def _check_tensor(t: torch.Tensor, what: str):
        if not isinstance(t, torch.Tensor):
            raise ValueError(f"{what} expects a torch.Tensor")
        if not t.is_contiguous():
            raise ValueError(f"{what} requires a contiguous tensor")
        n = t.numel() * t.element_size()
        if n == 0:
            raise ValueError(f"{what} requires a non-empty tensor")
        return n

    def _reg(self, kind: str, buffer_id: int, tensor: torch.Tensor, publish: bool) -> bool:
        if int(buffer_id) == 0:
            raise ValueError("buffer_id must be non-zero")
        n = self._check_tensor(tensor, f"reg_{kind}")
        ok, mr = self._ep.reg(tensor.data_ptr(), n)
        if not ok:
            return False
        with self._mu:
            self._bind[tensor.data_ptr()] = (tensor, int(buffer_id), n, mr, kind)
            self._ids[int(buffer_id)] = tensor.data_ptr()
        if publish:
            self._x.put(f"{kind}/{self._rank}/{int(buffer_id)}", n)
        return True

    def _unreg(self, kind: str, buffer_id: int) -> bool:
        with self._mu:
            ptr = self._ids.get(int(buffer_id))
            b = self._bind.get(ptr) if ptr is not None else None
            if b is None or (kind is not None and b[4] != kind):
                return False
            del self._ids[int(buffer_id)]
            del self._bind[ptr]
        self._x.delete(f"{b[4]}/{self._rank}/{int(buffer_id)}")
        return bool(self._ep.dereg(b[3]))

    def reg_rdma(self, buffer_id: int, tensor: torch.Tensor, publish: bool = True) -> bool:
        return self._reg("mr", buffer_id, tensor, publish)

    def unreg_rdma(self, buffer_id: int) -> bool:
        return self._unreg("mr", buffer_id)

    def reg_ipc(self, buffer_id: int, tensor: torch.Tensor, publish: bool = True) -> bool:
        return self._reg("ipc", buffer_id, tensor, publish)

    def unreg_ipc(self, buffer_id: int) -> bool:
        return self._unreg("ipc", buffer_id)

    def wait_mr(self, peer_rank: int, buffer_id: int, timeout_ms: int = 60000) -> bool:
        return bool(self._x.get(f"mr/{int(peer_rank)}/{int(buffer_id)}", timeout_ms)[0])

    def wait_ipc(self, peer_rank: int, buffer_id: int, timeout_ms: int = 60000) -> bool:
        return bool(self._x.get(f"ipc/{int(peer_rank)}/{int(buffer_id)}", timeout_ms)[0])

    # ------------------------------------------------------------------ transfers
    def _post(self, send: bool, peer_rank: int, tensor: torch.Tensor, offset: int, len: int) -> int:  # noqa: A002
        what = "isend" if send else "irecv"
        total = self._check_tensor(tensor, what)
        if len == 0:
            len = total - offset  # noqa: A001
        if offset < 0 or len <= 0 or offset + len > total:
            raise ValueError(f"{what} offset+len exceeds tensor size")
        conn = self._conn(peer_rank)
        with self._mu:
            b = self._bind.get(tensor.data_ptr())
        if b is not None and b[2] != total:
            raise RuntimeError("registered tensor size mismatch")
        tmp = None
        if b is None:
            ok, tmp = self._ep.reg(tensor.data_ptr(), total)
            if not ok:
                raise RuntimeError(f"{what} failed to register a temporary region")
        mr = b[3] if b is not None else tmp
        fn = self._ep.send_async if send else self._ep.recv_async
        ok, tid = fn(conn, mr, tensor.data_ptr() + offset, len)
        if not ok:
            if tmp is not None:
                self._ep.dereg(tmp)
            return 0
        with self._mu:
            req = self._next_req
            self._next_req += 1
            self._pending[req] = (tid, tensor, tmp)
        return req

    def isend(self, peer_rank: int, tensor: torch.Tensor, offset: int = 0, len: int = 0,  # noqa: A002
              remote_buffer_id: int = 0, remote_offset: int = 0) -> int:
        """Two-sided send of bytes [offset, offset + len) of `tensor` (len = 0: through the end).  In the reference
        `remote_buffer_id` / `remote_offset` are a placement hint for its zero-copy adapters; the matching `irecv`
        decides placement here, so the hint is only checked (the buffer must have been published)."""
        if remote_buffer_id:
            if not (self.wait_mr(peer_rank, remote_buffer_id, 0) or self.wait_ipc(peer_rank, remote_buffer_id, 0)):
                raise RuntimeError(f"rank {peer_rank} has not published buffer {remote_buffer_id}")
        return self._post(True, int(peer_rank), tensor, int(offset), int(len))

    def irecv(self, peer_rank: int, tensor: torch.Tensor, offset: int = 0, len: int = 0) -> int:  # noqa: A002
        return self._post(False, int(peer_rank), tensor, int(offset), int(len))

    def _finish(self, req: int):
        with self._mu:
            p = self._pending.pop(int(req), None)
        if p is not None and p[2] is not None:
            self._ep.dereg(p[2])

    def poll(self, req: int) -> bool:
        with self._mu:
            p = self._pending.get(int(req))
        if p is None:
            return True
        ok, done = self._ep.poll_async(p[0])
        if not ok:
            self._finish(req)
            raise RuntimeError(f"request {req} failed")
        if done:
            self._finish(req)
        return bool(done)

    def release(self, req: int) -> None:
        self._finish(req)

    def wait_finish(self, req: int) -> bool:
        with self._mu:
            p = self._pending.get(int(req))
        if p is None:
            return int(req) != 0
        try:
            return bool(self._ep.wait(p[0], -1))
        finally:
            self._finish(req)

    def wait_finish_multi(self, reqs) -> bool:
        ok = True
        for r in list(reqs):
            ok = self.wait_finish(r) and ok
        return ok

    def send(self, peer_rank: int, tensor: torch.Tensor, remote_buffer_id: int = 0, remote_offset: int = 0) -> None:
        req = self.isend(peer_rank, tensor, 0, 0, remote_buffer_id, remote_offset)
        if req == 0 or not self.wait_finish(req):
            raise RuntimeError(f"send to rank {peer_rank} failed")

    def recv(self, peer_rank: int, tensor: torch.Tensor) -> None:
        req = self.irecv(peer_rank, tensor)
        if req == 0 or not self.wait_finish(req):
            raise RuntimeError(f"recv from rank {peer_rank} failed")

    # ------------------------------------------------------------------ barrier / teardown
    def barrier(self, barrier_namespace: str = "default", timeout_ms: int = -1) -> bool:
        seq = self._barrier_seq.get(barrier_namespace, 0)
        self._barrier_seq[barrier_namespace] = seq + 1
        prefix = f"barrier/{barrier_namespace}/{seq}/"
        self._x.put(prefix + str(self._rank), True)
        return self._x.count(prefix, self._world, timeout_ms)

    def close(self) -> None:
        """Releases requests, registrations and the endpoint.  Rank 0 also stops the exchanger: have every rank pass a
        `barrier()` first."""
        if self._closed:
            return
        self._closed = True
        with self._mu:
            pend = list(self._pending)
            ids = list(self._ids)
        for r in pend:
            self._finish(r)
        for i in ids:
            self._unreg(None, i)
        self._x.close()
        if self._server is not None:
            self._server.close()
        self._ep = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass
