"""``deep_ep`` compatibility package: frameworks that ``import deep_ep`` (vLLM, SGLang, Megatron)
get the uccl_b200 implementation -- the role of the reference's ep/deep_ep_wrapper/deep_ep."""
from uccl_b200.ep import Buffer, Config, EventOverlap  # noqa: F401
from uccl_b200.ep.utils import EventHandle  # noqa: F401

__all__ = ["Buffer", "Config", "EventOverlap", "EventHandle"]
