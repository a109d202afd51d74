"""Backend / ReduceOp / option types of the collective API.

API parity with ray.util.collective.types (reference python/ray/util/collective/types.py:34-122):
the same public names (`Backend`, `ReduceOp`, `AllReduceOptions`, ... `RecvOptions`,
`unset_timeout_ms`), the same defaults, and option objects that are plain attribute bags.  Added:
the `B200` backend constant (with `NCCL` kept as an alias so actor code written for the
reference's default backend runs unchanged) and `DagReduceOp`, the second ReduceOp numbering that
compiled graphs pass raw to the communicator (python/ray/experimental/util/types.py:12-17).
"""
import enum
from datetime import timedelta


def _probe(module_name):
    try:
        return __import__(module_name)
    except ImportError:
        return None


th = _probe("torch")
cp = _probe("cupy")


def torch_available() -> bool:
    return th is not None


def cupy_available() -> bool:
    return cp is not None


class Backend(object):
    """`Backend("nccl")` normalises a user string to one of the constants below."""

    B200 = "B200"
    NCCL = B200  # drop-in: the reference's default backend name selects the B200 kernels
    GLOO = "GLOO"
    UNRECOGNIZED = "unrecognized"
    _ALIASES = {"TORCH_GLOO": "GLOO"}

    def __new__(cls, name: str):
        key = cls._ALIASES.get(name.upper(), name.upper())
        value = getattr(Backend, key, Backend.UNRECOGNIZED) if not key.startswith("_") else Backend.UNRECOGNIZED
        if value == Backend.UNRECOGNIZED:
            raise ValueError("Unrecognized backend: '{}'. Only B200 (alias NCCL) and GLOO are supported".format(name))
        return value


class ReduceOp(enum.Enum):
    """ray.util.collective numbering."""

    SUM = 0
    PRODUCT = 1
    MIN = 2
    MAX = 3


class DagReduceOp(enum.Enum):
    """ray.experimental.util.types.ReduceOp numbering, identical to ncclRedOp_t."""

    SUM = 0
    PRODUCT = 1
    MAX = 2
    MIN = 3
    AVG = 4


unset_timeout_ms = timedelta(milliseconds=-1)


def _option_type(name: str, doc: str, **defaults):
    """An attribute bag with class-level defaults; instances are created empty and callers assign
    the fields they need, which is how the reference's API layer uses its option dataclasses."""
    body = dict(defaults)
    body["__doc__"] = doc
    body["__repr__"] = lambda self: "{}({})".format(
        name, ", ".join("{}={!r}".format(k, getattr(self, k)) for k in defaults))
    return type(name, (object,), body)


AllReduceOptions = _option_type("AllReduceOptions", "allreduce: reduction operator", reduceOp=ReduceOp.SUM,
                                timeout_ms=unset_timeout_ms)
BarrierOptions = _option_type("BarrierOptions", "barrier", timeout_ms=unset_timeout_ms)
ReduceOptions = _option_type("ReduceOptions", "reduce: operator, destination rank (and tensor index)",
                             reduceOp=ReduceOp.SUM, root_rank=0, root_tensor=0, timeout_ms=unset_timeout_ms)
AllGatherOptions = _option_type("AllGatherOptions", "allgather", timeout_ms=unset_timeout_ms)
BroadcastOptions = _option_type("BroadcastOptions", "broadcast: source rank (and tensor index)", root_rank=0,
                                root_tensor=0, timeout_ms=unset_timeout_ms)
ReduceScatterOptions = _option_type("ReduceScatterOptions", "reducescatter: reduction operator",
                                    reduceOp=ReduceOp.SUM, timeout_ms=unset_timeout_ms)
SendOptions = _option_type("SendOptions", "send: destination rank, optional element-count prefix", dst_rank=0,
                           dst_gpu_index=0, n_elements=0, timeout_ms=unset_timeout_ms)
RecvOptions = _option_type("RecvOptions", "recv: source rank, optional element-count prefix", src_rank=0,
                           src_gpu_index=0, n_elements=0, timeout_ms=unset_timeout_ms)
