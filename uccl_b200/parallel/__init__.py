from .comm import ALGOS, Communicator, dtype_code, op_code  # noqa: F401
