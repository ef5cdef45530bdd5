"""``deep_ep.buffer`` of upstream DeepEP: ``from deep_ep.buffer import Buffer`` resolves to the same class as
``deep_ep.Buffer``."""
from . import Buffer  # noqa: F401
from uccl_b200.ep import Config  # noqa: F401
from uccl_b200.ep.utils import EventHandle, EventOverlap  # noqa: F401

__all__ = ["Buffer", "Config", "EventOverlap", "EventHandle"]
