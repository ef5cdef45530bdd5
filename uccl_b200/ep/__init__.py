"""Expert-parallel dispatch/combine (DeepEP API). Mirrors ``uccl.ep`` + ``ep/bench/buffer.py``."""
from .buffer import Buffer, Config  # noqa: F401
from .utils import (EventHandle, EventOverlap, bench, calc_diff, inplace_unique,  # noqa: F401
                    per_token_cast_back, per_token_cast_to_fp8, pack_ue8m0, unpack_ue8m0, hash_tensor,
                    create_grouped_scores, init_dist, detect_group_topology, check_nvlink_connections,
                    initialize_uccl, destroy_uccl, bench_kineto, detect_ib_hca, get_peer_ip,
                    get_cpu_proxies_meta, logfmt10_simulate)
from .proxy import FifoProxy, Proxy  # noqa: F401,E402
from .autograd import ep_combine, ep_dispatch  # noqa: F401,E402


# ---------------------------------------------------------------------------------------------------------------
# Module-level functions of the reference's native module (`uccl.ep`, ep/src/uccl_ep.cc:1641-2410) that scripts
# written against it call directly.
def get_low_latency_rdma_size_hint(num_max_dispatch_tokens_per_rank: int, hidden: int, num_ranks: int,
                                   num_experts: int) -> int:
    return Buffer.get_low_latency_rdma_size_hint(num_max_dispatch_tokens_per_rank, hidden, num_ranks, num_experts)


def is_sm90_compiled() -> bool:
    return Buffer.is_sm90_compiled()


def get_oob_ip() -> str:
    from ..p2p import get_oob_ip as _ip

    return _ip()


def get_num_proxy_threads() -> int:
    """The reference runs 4 proxy threads x 8 FIFOs per GPU to feed its NICs (ep/include/common.hpp); a
    :class:`Proxy` here is one service thread per communicator (peers are load/store reachable)."""
    return 1


_PROXIES_BY_DEVICE = {}


def register_proxies(device_index: int, proxies) -> None:
    """Remember the proxies serving a device so that :func:`stop_all_registered_proxies` can stop them
    (reference: ep/src/uccl_ep.cc:1676-1712, used by ep/bench/utils.py:602)."""
    _PROXIES_BY_DEVICE.setdefault(int(device_index), []).extend(list(proxies))


def stop_all_registered_proxies() -> None:
    for plist in _PROXIES_BY_DEVICE.values():
        for p in plist:
            try:
                p.stop()
            except Exception:  # a proxy that is already down must not keep the others running
                pass
    _PROXIES_BY_DEVICE.clear()


def can_register_rdma_gpu_buffer(device_index: int, num_bytes: int) -> bool:
    """Always true: the buffers peers touch live in the VMM symmetric heap, nothing is registered with a NIC."""
    return True


def rdma_buffer_should_use_host_alloc(device_index: int, num_bytes: int = 4096) -> bool:
    return False


def get_rdma_buffer(num_rdma_bytes: int, device_index: int):
    """``(tensor, is_host_allocated)`` like the reference's DLPack scratch allocation (ep/src/uccl_ep.cc:1723-1733).
    Kept for source compatibility: ``Buffer`` places its low-latency block in the symmetric heap itself and does not
    take an external scratch tensor, so the returned device tensor is ordinary memory."""
    import torch

    dev = torch.device("cuda", int(device_index)) if torch.cuda.is_available() and device_index >= 0 else torch.device("cpu")
    return torch.zeros(int(num_rdma_bytes), dtype=torch.uint8, device=dev), False
