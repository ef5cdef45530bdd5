"""``deep_ep.utils`` of upstream DeepEP (deep_ep/utils.py): ``EventOverlap`` and ``check_nvlink_connections`` -- some
consumers import them from here rather than from the package root."""
from uccl_b200.ep.utils import EventHandle, EventOverlap, check_nvlink_connections  # noqa: F401

__all__ = ["EventOverlap", "EventHandle", "check_nvlink_connections"]
