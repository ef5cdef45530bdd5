"""``deep_ep`` compatibility package: frameworks that ``import deep_ep`` (vLLM, SGLang, Megatron)
get the uccl_b200 implementation -- the role of the reference's ep/deep_ep_wrapper/deep_ep.

Code written against DeepEP keeps ``recv_x`` and friends for as long as it likes (saved activations, several MoE
layers in flight), so the ``Buffer`` exported here hands out OWNED tensors like DeepEP does
(``owned_results=True``); ``uccl_b200.ep.Buffer`` itself defaults to zero-copy views of its receive arenas."""
from uccl_b200.ep import Buffer as _Buffer
from uccl_b200.ep import Config, EventOverlap  # noqa: F401
from uccl_b200.ep.utils import EventHandle, check_nvlink_connections, destroy_uccl, initialize_uccl  # noqa: F401


class Buffer(_Buffer):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("owned_results", True)
        super().__init__(*args, **kwargs)


# the reference's wrapper also exports its set-up helpers from the package root (ep/deep_ep_wrapper/deep_ep/__init__.py)
__all__ = ["Buffer", "Config", "EventOverlap", "EventHandle", "check_nvlink_connections", "initialize_uccl", "destroy_uccl"]
