"""Name of upstream DeepEP's C++ extension (``import deep_ep_cpp; deep_ep_cpp.Config(...)``): code that reaches for it
gets this library's ``Config`` and ``EventHandle``.  The extension's ``Buffer`` (the runtime object behind
``deep_ep.Buffer.runtime``) is an implementation detail of upstream and is not reproduced."""
from uccl_b200.ep import Config  # noqa: F401
from uccl_b200.ep.utils import EventHandle  # noqa: F401

__all__ = ["Config", "EventHandle"]
