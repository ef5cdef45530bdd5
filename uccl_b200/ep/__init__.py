"""Expert-parallel dispatch/combine (DeepEP API). Mirrors ``uccl.ep`` + ``ep/bench/buffer.py``."""
from .buffer import Buffer, Config  # noqa: F401
from .utils import (EventHandle, EventOverlap, bench, calc_diff, inplace_unique,  # noqa: F401
                    per_token_cast_back, per_token_cast_to_fp8, pack_ue8m0, unpack_ue8m0, hash_tensor,
                    create_grouped_scores, init_dist, detect_group_topology, check_nvlink_connections,
                    initialize_uccl, destroy_uccl, bench_kineto)
from .proxy import FifoProxy, Proxy  # noqa: F401,E402
from .autograd import ep_combine, ep_dispatch  # noqa: F401,E402
