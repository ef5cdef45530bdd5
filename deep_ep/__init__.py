"""``deep_ep`` compatibility package: frameworks that ``import deep_ep`` (vLLM, SGLang, Megatron)
get the uccl_b200 implementation -- the role of the reference's ep/deep_ep_wrapper/deep_ep.

Code written against DeepEP keeps ``recv_x`` and friends for as long as it likes (saved activations, several MoE
layers in flight), so the ``Buffer`` exported here hands out OWNED tensors like DeepEP does
(``owned_results=True``); ``uccl_b200.ep.Buffer`` itself defaults to zero-copy views of its receive arenas."""
from uccl_b200.ep import Buffer as _Buffer
from uccl_b200.ep import Config, EventOverlap  # noqa: F401
from uccl_b200.ep.utils import EventHandle  # noqa: F401


class Buffer(_Buffer):
    def __init__(self, *args, **kwargs):
        kwargs.setdefault("owned_results", True)
        super().__init__(*args, **kwargs)


__all__ = ["Buffer", "Config", "EventOverlap", "EventHandle"]
