from .comm import ALGOS, Communicator, dtype_code, op_code  # noqa: F401


def __getattr__(name):  # lazy: the network stack is only needed for multi-node jobs
    if name in ("MultiNodeCommunicator", "NativeMultiNodeCommunicator", "AsyncMultiNode", "MultiNodeWork"):
        from . import multinode

        return getattr(multinode, name)
    raise AttributeError(name)
