"""Backend / ReduceOp / option types of the collective API.

Mirrors ray.util.collective.types (reference python/ray/util/collective/types.py:34-122):
same names, same defaults, same "options are classes with class attributes" shape, plus the
`B200` backend constant that selects the peer-memory kernels.  `NCCL` is accepted as an alias
of `B200` so that actor code written for the reference's default backend runs unchanged.
The second ReduceOp numbering used by compiled graphs
(python/ray/experimental/util/types.py:12-17) is `DagReduceOp`.
"""
from dataclasses import dataclass
from datetime import timedelta
from enum import Enum

try:
    import torch as th  # noqa: F401

    _TORCH_AVAILABLE = True
except ImportError:  # pragma: no cover
    _TORCH_AVAILABLE = False

try:
    import cupy as cp  # noqa: F401

    _CUPY_AVAILABLE = True
except ImportError:
    _CUPY_AVAILABLE = False


def cupy_available():
    return _CUPY_AVAILABLE


def torch_available():
    return _TORCH_AVAILABLE


class Backend(object):
    """String-enum of backends; `Backend("nccl")` etc. normalises a user string."""

    B200 = "B200"
    NCCL = "B200"  # drop-in: the reference's default backend name selects the B200 kernels
    GLOO = "GLOO"
    UNRECOGNIZED = "unrecognized"

    def __new__(cls, name: str):
        upper_name = name.upper()
        backend = getattr(Backend, upper_name, Backend.UNRECOGNIZED)
        if backend == Backend.UNRECOGNIZED:
            if upper_name == "TORCH_GLOO":
                return Backend.GLOO
            raise ValueError(
                "Unrecognized backend: '{}'. Only B200 (alias NCCL) and GLOO are supported".format(name))
        return backend


class ReduceOp(Enum):
    """ray.util.collective numbering (types.py:55-59)."""

    SUM = 0
    PRODUCT = 1
    MIN = 2
    MAX = 3


class DagReduceOp(Enum):
    """ray.experimental.util.types.ReduceOp numbering == ncclRedOp_t (util/types.py:12-17)."""

    SUM = 0
    PRODUCT = 1
    MAX = 2
    MIN = 3
    AVG = 4


unset_timeout_ms = timedelta(milliseconds=-1)


@dataclass
class AllReduceOptions:
    reduceOp = ReduceOp.SUM
    timeout_ms = unset_timeout_ms


@dataclass
class BarrierOptions:
    timeout_ms = unset_timeout_ms


@dataclass
class ReduceOptions:
    reduceOp = ReduceOp.SUM
    root_rank = 0
    root_tensor = 0
    timeout_ms = unset_timeout_ms


@dataclass
class AllGatherOptions:
    timeout_ms = unset_timeout_ms


@dataclass
class BroadcastOptions:
    root_rank = 0
    root_tensor = 0
    timeout_ms = unset_timeout_ms


@dataclass
class ReduceScatterOptions:
    reduceOp = ReduceOp.SUM
    timeout_ms = unset_timeout_ms


@dataclass
class SendOptions:
    dst_rank = 0
    dst_gpu_index = 0
    n_elements = 0
    timeout_ms = unset_timeout_ms


@dataclass
class RecvOptions:
    src_rank = 0
    src_gpu_index = 0
    n_elements = 0
    unset_timeout_ms = unset_timeout_ms
