"""NIXL-style P2P transfer engine (KV-cache / weight moves between GPUs of one node).

Python surface mirrors the reference's ``uccl.p2p.Endpoint`` (p2p/engine_api.cc:141-1517;
list in SURVEY 2.6) so callers switch over unchanged; the engine underneath is the native
``P2PEndpoint`` (csrc/p2p): peer HBM mapped once per allocation through CUDA IPC, bytes moved
by an in-kernel TMA (cp.async.bulk) copy pipeline on side streams, vectorised per launch.
"""
from __future__ import annotations

import enum
import os
import socket
import struct
from typing import List, NamedTuple, Optional, Sequence, Tuple

from .. import _native

XFER_DESC_BYTES = 128


class XferDesc:
    """A registered memory window (reference: XferDesc, p2p/engine_api.cc:15-22)."""

    __slots__ = ("raw", "mr_id")

    def __init__(self, raw: bytes, mr_id: int = 0):
        assert len(raw) == XFER_DESC_BYTES
        self.raw = bytes(raw)
        self.mr_id = mr_id

    @property
    def addr(self) -> int:
        return struct.unpack_from("<Q", self.raw, 72)[0]

    @property
    def size(self) -> int:
        return struct.unpack_from("<Q", self.raw, 80)[0]

    @property
    def lkeys(self) -> list:
        """NIC keys of the reference's descriptor (p2p/engine_api.cc:163-164).  Always empty: nothing is registered with
        a NIC, the window is reached by load/store (the reference calls this "IPC-only metadata")."""
        return []

    @property
    def rkeys(self) -> list:
        return []

    def __repr__(self):
        return f"XferDesc(addr=0x{self.addr:x}, size={self.size}, mr_id={self.mr_id})"


class FloatType(enum.IntEnum):
    """Element type a registration may carry (reference: uccl::FloatType, p2p/engine_api.cc:171-177); the lossless
    compression hook (`uccl_b200.p2p.compress`) uses it to pick the exponent / mantissa split."""

    kUndefined = 0
    kFloat16 = 1
    kBFloat16 = 2
    kFloat32 = 3
    kFloat8E4M3FN = 4
    kFloat8E5M2 = 5

    @staticmethod
    def from_tensor(t) -> "FloatType":
        import torch

        return {torch.float16: FloatType.kFloat16, torch.bfloat16: FloatType.kBFloat16, torch.float32: FloatType.kFloat32,
                torch.float8_e4m3fn: FloatType.kFloat8E4M3FN, torch.float8_e5m2: FloatType.kFloat8E5M2}.get(
                    t.dtype, FloatType.kUndefined)


class XferHandle(NamedTuple):
    """Identifies a started transfer (reference: XferHandle, p2p/engine_api.cc:24,147-150)."""

    conn_id: int
    op_name: str
    transfer_id: int


def get_oob_ip() -> str:
    return os.environ.get("UCCL_B200_P2P_IP", "127.0.0.1")


def _registry_path(gpu_idx: int) -> str:
    return f"/dev/shm/uccl_b200_p2p_{os.getuid()}_gpu{gpu_idx}"


class Endpoint:
    def __init__(self, local_gpu_idx: Optional[int] = None, num_cpus: int = 4):
        import torch

        C = _native.C()
        if local_gpu_idx is None:
            local_gpu_idx = torch.cuda.current_device() if torch.cuda.is_available() else -1
        if int(local_gpu_idx) >= 0 and not torch.cuda.is_available() and os.environ.get("UCCL_B200_P2P_HOST_FALLBACK") == "1":
            # GPU-less CI running scripts that name a GPU: same engine in host mode (buffers must then be host memory)
            local_gpu_idx = -1
        if int(local_gpu_idx) >= 0:
            torch.cuda.init()
        # local_gpu_idx < 0: host mode (buffers are host memory of this process, copies are memcpy) --
        # the control plane, matching and transfer bookkeeping are the production code (GPU-less CI)
        self._e = C.P2PEndpoint(int(local_gpu_idx), max(1, int(num_cpus)))
        self.local_gpu_idx = int(local_gpu_idx)
        self._mrs = {}
        self._mr_types = {}
        self._rank2conn = {}
        self._ipc_pending = {}
        # local rendezvous file for connect_local (reference: shm inbox keyed by GPU BDF)
        try:
            with open(_registry_path(self.local_gpu_idx), "wb") as f:
                f.write(self.get_metadata())
        except OSError:
            pass

    # ------------------------------------------------------------------ metadata
    def get_metadata(self) -> bytes:
        return bytes(self._e.get_metadata())

    @staticmethod
    def parse_metadata(metadata: bytes) -> Tuple[str, int, int]:
        return _native.C().P2PEndpoint.parse_metadata(bytes(metadata))

    # --------------------------------------------------------------- connections
    def connect(self, remote_ip_addr=None, remote_gpu_idx: int = 0, remote_port: int = 0, remote_metadata=None,
                remote_gpu_bdf=None, ip_addr=None):
        """``connect(ip, gpu, port)`` like the reference (whose second argument is the peer GPU's PCI BDF; here the GPU
        index -- `remote_gpu_bdf` is accepted as its keyword), or ``connect(remote_metadata=md)``."""
        if remote_ip_addr is None:
            remote_ip_addr = ip_addr
        if remote_gpu_bdf is not None:
            remote_gpu_idx = int(remote_gpu_bdf)
        if remote_metadata is None and isinstance(remote_ip_addr, (bytes, bytearray)):
            remote_metadata = remote_ip_addr
        if remote_metadata is not None:
            return self._e.add_remote_endpoint(bytes(remote_metadata))
        return self._e.connect(str(remote_ip_addr), int(remote_gpu_idx), int(remote_port))

    def set_rank_conn(self, rank: int, conn_id: int) -> None:
        """Remember which connection leads to a peer rank (filled by `uccl_b200.collective`)."""
        self._rank2conn[int(rank)] = int(conn_id)

    def conn_id_of_rank(self, rank: int) -> int:
        """Connection id of a peer rank, 2**64 - 1 if none (reference: Endpoint::conn_id_of_rank, p2p/engine.h:456)."""
        return self._rank2conn.get(int(rank), 0xFFFFFFFFFFFFFFFF)

    def accept(self, timeout_ms: int = -1):
        return self._e.accept(timeout_ms)

    def start_passive_accept(self) -> bool:
        return self._e.start_passive_accept()

    def add_remote_endpoint(self, metadata_bytes: bytes):
        return self._e.add_remote_endpoint(bytes(metadata_bytes))

    def remove_remote_endpoint(self, conn_id: int) -> bool:
        return self._e.remove_remote_endpoint(conn_id)

    def connect_local(self, remote_gpu_idx: Optional[int] = None, remote_gpu_bdf=None):
        if remote_gpu_idx is None:
            remote_gpu_idx = remote_gpu_bdf
        with open(_registry_path(int(remote_gpu_idx)), "rb") as f:
            md = f.read()
        return self._e.add_remote_endpoint(md)

    def accept_local(self, timeout_ms: int = -1):
        ok, ip, gpu, conn = self._e.accept(timeout_ms)
        return ok, gpu, conn

    # -------------------------------------------------------------- registration
    def reg(self, ptr: int, size: int, floatType: FloatType = FloatType.kUndefined):
        ok, mr = self._e.reg(int(ptr), int(size))
        if ok:
            self._mrs[mr] = (int(ptr), int(size))
            self._mr_types[mr] = FloatType(int(floatType))
        return ok, mr

    def float_type(self, mr_id: int) -> FloatType:
        """Element type recorded at registration (kUndefined if none was given)."""
        return self._mr_types.get(mr_id, FloatType.kUndefined)

    def __repr__(self):
        mode = f"gpu {self.local_gpu_idx}" if self.local_gpu_idx >= 0 else "host mode"
        return f"<uccl_b200 P2P Endpoint ({mode}, {len(self._mrs)} registrations)>"

    def regv(self, ptrs: Sequence[int], sizes: Sequence[int]):
        ids = []
        for p, s in zip(ptrs, sizes):
            ok, mr = self.reg(p, s)
            if not ok:
                return False, ids
            ids.append(mr)
        return True, ids

    def dereg(self, mr_id: int) -> bool:
        self._mrs.pop(mr_id, None)
        self._mr_types.pop(mr_id, None)
        return self._e.dereg(mr_id)

    def register_memory(self, tensor_list) -> List[XferDesc]:
        descs = []
        for t in tensor_list:
            ptr, size = t.data_ptr(), t.numel() * t.element_size()
            ok, mr = self.reg(ptr, size, FloatType.from_tensor(t))
            if not ok:
                raise RuntimeError("uccl_b200.p2p: register_memory failed")
            descs.append(XferDesc(bytes(self._e.describe(ptr, size)), mr))
        return descs

    def deregister_memory(self, desc_list: Sequence[XferDesc]) -> None:
        for d in desc_list:
            if d.mr_id:
                self.dereg(d.mr_id)

    @staticmethod
    def get_serialized_descs(desc_list: Sequence[XferDesc]) -> bytes:
        return struct.pack("<I", len(desc_list)) + b"".join(d.raw for d in desc_list)

    @staticmethod
    def deserialize_descs(serialized_bytes: bytes) -> List[XferDesc]:
        (n,) = struct.unpack_from("<I", serialized_bytes, 0)
        return [XferDesc(serialized_bytes[4 + i * XFER_DESC_BYTES: 4 + (i + 1) * XFER_DESC_BYTES]) for i in range(n)]

    # ------------------------------------------------------------------ two-sided
    def send_async(self, conn_id, mr_id, ptr, size):
        return self._e.send_async(conn_id, [int(ptr)], [int(size)])

    def recv_async(self, conn_id, mr_id, ptr, size):
        return self._e.recv_async(conn_id, [int(ptr)], [int(size)])

    def sendv_async(self, conn_id, mr_id_v, data_ptr_v, size_v, num_iovs=None):
        return self._e.send_async(conn_id, [int(p) for p in data_ptr_v], [int(s) for s in size_v])

    def recvv_async(self, conn_id, mr_id_v, data_ptr_v, size_v, num_iovs=None):
        return self._e.recv_async(conn_id, [int(p) for p in data_ptr_v], [int(s) for s in size_v])

    def _block(self, res) -> bool:
        ok, tid = res
        return bool(ok) and bool(self._e.wait(tid, -1))

    def send(self, conn_id, mr_id, ptr, size) -> bool:
        return self._block(self.send_async(conn_id, mr_id, ptr, size))

    def recv(self, conn_id, mr_id, ptr, size) -> bool:
        return self._block(self.recv_async(conn_id, mr_id, ptr, size))

    # ---- compression hook (reference: UCCL_P2P_COMPRESS_STRATEGY, p2p/rdma/compression.h) -- a 16-byte
    # header {compressed?, payload bytes} precedes the payload; both sides must use the *_compressed pair
    def send_compressed(self, conn_id, tensor, strategy=None) -> bool:
        import struct

        import torch

        from .compress import Compressor

        comp = Compressor(strategy)
        raw_bytes = tensor.numel() * tensor.element_size()
        if comp.wants(tensor):
            payload, nbytes = comp.compress(tensor)
            flag = 1
        else:
            payload, nbytes, flag = tensor, raw_bytes, 0
        hdr = torch.frombuffer(bytearray(struct.pack("<QQ", flag, nbytes)), dtype=torch.uint8).to(tensor.device)
        torch.cuda.current_stream(tensor.device).synchronize()
        return self.send(conn_id, 0, hdr.data_ptr(), 16) and self.send(conn_id, 0, payload.data_ptr(), nbytes)

    def recv_compressed(self, conn_id, out) -> bool:
        import struct

        import torch

        from .compress import Compressor

        hdr = torch.zeros(16, dtype=torch.uint8, device=out.device)
        torch.cuda.current_stream(out.device).synchronize()
        if not self.recv(conn_id, 0, hdr.data_ptr(), 16):
            return False
        flag, nbytes = struct.unpack("<QQ", bytes(hdr.cpu().numpy().tobytes()))
        if not flag:
            assert nbytes == out.numel() * out.element_size()
            return self.recv(conn_id, 0, out.data_ptr(), nbytes)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=out.device)
        torch.cuda.current_stream(out.device).synchronize()
        if not self.recv(conn_id, 0, buf.data_ptr(), nbytes):
            return False
        Compressor("for").decompress(buf, out.numel(), out.dtype, out=out.view(-1))
        torch.cuda.current_stream(out.device).synchronize()
        return True

    def sendv(self, conn_id, mr_id_v, data_ptr_v, size_v, num_iovs=None) -> bool:
        return self._block(self.sendv_async(conn_id, mr_id_v, data_ptr_v, size_v))

    def recvv(self, conn_id, mr_id_v, data_ptr_v, size_v, num_iovs=None) -> bool:
        return self._block(self.recvv_async(conn_id, mr_id_v, data_ptr_v, size_v))

    # ------------------------------------------------------------------ one-sided
    def advertise(self, conn_id, mr_id, ptr, size):
        return True, bytes(self._e.describe(int(ptr), int(size)))

    def advertisev(self, conn_id, mr_id_v, ptr_v, size_v, num_iovs=None):
        return True, [bytes(self._e.describe(int(p), int(s))) for p, s in zip(ptr_v, size_v)]

    @staticmethod
    def _raw(d):
        return d.raw if isinstance(d, XferDesc) else bytes(d)

    def write_async(self, conn_id, mr_id, ptr, size, meta):
        return self._e.write_async(conn_id, [int(ptr)], [int(size)], [self._raw(meta)])

    def read_async(self, conn_id, mr_id, ptr, size, meta):
        return self._e.read_async(conn_id, [int(ptr)], [int(size)], [self._raw(meta)])

    def writev_async(self, conn_id, mr_id_v, ptr_v, size_v, meta_blob_v, num_iovs=None):
        return self._e.write_async(conn_id, [int(p) for p in ptr_v], [int(s) for s in size_v],
                                   [self._raw(b) for b in meta_blob_v])

    def readv_async(self, conn_id, mr_id_v, ptr_v, size_v, meta_blob_v, num_iovs=None):
        return self._e.read_async(conn_id, [int(p) for p in ptr_v], [int(s) for s in size_v],
                                  [self._raw(b) for b in meta_blob_v])

    def write(self, conn_id, mr_id, ptr, size, meta) -> bool:
        return self._block(self.write_async(conn_id, mr_id, ptr, size, meta))

    def read(self, conn_id, mr_id, ptr, size, meta) -> bool:
        return self._block(self.read_async(conn_id, mr_id, ptr, size, meta))

    def writev(self, conn_id, mr_id_v, ptr_v, size_v, meta_blob_v, num_iovs=None) -> bool:
        return self._block(self.writev_async(conn_id, mr_id_v, ptr_v, size_v, meta_blob_v))

    def readv(self, conn_id, mr_id_v, ptr_v, size_v, meta_blob_v, num_iovs=None) -> bool:
        return self._block(self.readv_async(conn_id, mr_id_v, ptr_v, size_v, meta_blob_v))

    # ------------------------------------------ IPC-named variants (same engine here)
    def send_ipc(self, conn_id, ptr, size) -> bool:
        return self.send(conn_id, 0, ptr, size)

    def recv_ipc(self, conn_id, ptr, size) -> bool:
        return self.recv(conn_id, 0, ptr, size)

    def send_ipc_async(self, conn_id, ptr, size):
        return self.send_async(conn_id, 0, ptr, size)

    def recv_ipc_async(self, conn_id, ptr, size):
        return self.recv_async(conn_id, 0, ptr, size)

    def advertise_ipc(self, conn_id, ptr, size):
        return self.advertise(conn_id, 0, ptr, size)

    def advertisev_ipc(self, conn_id, ptr_v, size_v, num_iovs=None):
        return self.advertisev(conn_id, None, ptr_v, size_v)

    def write_ipc(self, conn_id, ptr, size, info) -> bool:
        return self.write(conn_id, 0, ptr, size, info)

    def read_ipc(self, conn_id, ptr, size, info) -> bool:
        return self.read(conn_id, 0, ptr, size, info)

    def write_ipc_async(self, conn_id, ptr, size, info):
        return self.write_async(conn_id, 0, ptr, size, info)

    def read_ipc_async(self, conn_id, ptr, size, info):
        return self.read_async(conn_id, 0, ptr, size, info)

    def writev_ipc(self, conn_id, ptr_v, size_v, info_v, num_iovs=None) -> bool:
        return self.writev(conn_id, None, ptr_v, size_v, info_v)

    def readv_ipc(self, conn_id, ptr_v, size_v, info_v, num_iovs=None) -> bool:
        return self.readv(conn_id, None, ptr_v, size_v, info_v)

    def writev_ipc_async(self, conn_id, ptr_v, size_v, info_v, num_iovs=None):
        return self.writev_async(conn_id, None, ptr_v, size_v, info_v)

    def readv_ipc_async(self, conn_id, ptr_v, size_v, info_v, num_iovs=None):
        return self.readv_async(conn_id, None, ptr_v, size_v, info_v)

    # ---------------------------------------------------------------- descriptor API
    def transfer(self, conn_id: int, op_name: str, local_desc_list: Sequence[XferDesc],
                 remote_desc_list: Sequence[XferDesc]):
        """NIXL-style: move every (local, remote) window pair with ONE kernel launch.  Parameter names as in the
        reference's binding (p2p/engine_api.cc:447-663)."""
        assert op_name in ("read", "write") and len(local_desc_list) == len(remote_desc_list)
        ptrs = [d.addr for d in local_desc_list]
        sizes = [min(l.size, r.size) for l, r in zip(local_desc_list, remote_desc_list)]
        blobs = [r.raw for r in remote_desc_list]
        fn = self._e.write_async if op_name == "write" else self._e.read_async
        return fn(conn_id, ptrs, sizes, blobs)

    # ---- prepared transfers (NIXL prepXfer / postXfer)
    def prepare_transfer(self, conn_id: int, op_name: str, local_desc_list: Sequence[XferDesc],
                         remote_desc_list: Sequence[XferDesc]):
        """Resolve the descriptor lists once (peer mapping, kernel descriptor tables in pinned memory).  Returns a
        handle for :meth:`post_transfer`; every post is then a bare kernel launch, however many blocks it moves --
        the shape of a KV-cache mover that re-sends the same page lists.  Raises if the peer is not load/store
        reachable (another host): use :meth:`transfer` there."""
        assert op_name in ("read", "write") and len(local_desc_list) == len(remote_desc_list)
        ptrs = [d.addr for d in local_desc_list]
        sizes = [min(l.size, r.size) for l, r in zip(local_desc_list, remote_desc_list)]
        ok, prep = self._e.prepare(conn_id, op_name == "write", ptrs, sizes, [r.raw for r in remote_desc_list])
        if not ok:
            raise RuntimeError("uccl_b200.p2p: prepare_transfer needs a load/store reachable peer and device memory")
        return prep

    def post_transfer(self, prep: int):
        """Launch a prepared transfer: ``(ok, transfer_id)`` like :meth:`transfer`."""
        return self._e.post(prep)

    def release_transfer(self, prep: int) -> bool:
        return self._e.release(prep)

    def poll_async(self, transfer_id: int):
        return self._e.poll_async(transfer_id)

    def wait(self, transfer_id: int, timeout_ms: int = -1) -> bool:
        return self._e.wait(transfer_id, timeout_ms)

    # -------------------------------------------------------------- notifications
    def send_notif(self, conn_id: int, msg: bytes) -> bool:
        return self._e.send_notif(conn_id, bytes(msg))

    def get_notifs(self):
        return self._e.get_notifs()

    def stats(self) -> dict:
        return dict(self._e.stats())


def __getattr__(name):  # inter-node channel: pulls in the network stack only when used
    if name == "NetChannel":
        from .internode import NetChannel

        return NetChannel
    raise AttributeError(name)
