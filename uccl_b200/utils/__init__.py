"""Utilities: device-timed benchmarking, fp8 helpers, the measured allreduce tuner, logging."""
from ..ep.utils import (bench, calc_diff, per_token_cast_back, per_token_cast_to_fp8)  # noqa: F401
from .metrics import MetricsExporter  # noqa: F401
from .sm_partition import SmPartition, sm_ids  # noqa: F401
from .tuner import autotune_allreduce, load_tuning, load_tuning_from_env, save_tuning, tuning_from_sweep  # noqa: F401


def set_log_level(level: str) -> None:
    """FATAL / ERROR / WARN / INFO / TRACE (same names as UCCL_DEBUG of the reference)."""
    from .. import _native

    _native.C().set_log_level({"FATAL": 0, "ERROR": 1, "WARN": 2, "INFO": 3, "TRACE": 4}[level.upper()])
