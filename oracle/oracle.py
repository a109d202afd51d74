"""Python face of the C oracle (oracle_reduce.c) plus the data-movement ops.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs; the product package never imports it.

Inputs and outputs are CPU torch tensors (torch is used only as a typed byte container so that
bf16 has a dtype; all arithmetic happens in the C library).  Semantics restated from the
reference call sites:
  allreduce      nccl_collective_group.py:168-190 / torch_gloo_collective_group.py:147-156
  reduce         :212-236 / :158-179 (result only defined on root)
  broadcast      :238-262 / :193-197
  allgather      :264-300 / :181-191  (rank j's tensor lands in slot j on every rank)
  reducescatter  :302-341 / :199-221  (rank r receives the fold of every rank's list[r])
  fused gradient mean (K13): torch's default / bf16_compress DDP hooks
      (torch/distributed/algorithms/ddp_comm_hooks/default_hooks.py:35-134): mean over ranks,
      optionally with every contribution and the result rounded to a 16-bit wire type.
"""
import ctypes
import os
from typing import List, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None

# ncclDataType_t numbering, same as include/b200coll.h
DT = {torch.int8: 0, torch.uint8: 1, torch.bool: 1, torch.int32: 2, torch.uint32: 3, torch.int64: 4, torch.uint64: 5,
      torch.float16: 6, torch.float32: 7, torch.float64: 8, torch.bfloat16: 9}
SUM, PROD, MAX, MIN, AVG = range(5)


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(path)
        _lib.oracle_reduce.restype = ctypes.c_int
        _lib.oracle_reduce.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t,
                                       ctypes.POINTER(ctypes.c_void_p), ctypes.c_void_p, ctypes.c_int, ctypes.c_float]
        for name in ("oracle_float_to_bf16", "oracle_float_to_f16"):
            getattr(_lib, name).restype = ctypes.c_uint16
            getattr(_lib, name).argtypes = [ctypes.c_float]
        for name in ("oracle_bf16_to_float", "oracle_f16_to_float"):
            getattr(_lib, name).restype = ctypes.c_float
            getattr(_lib, name).argtypes = [ctypes.c_uint16]
    return _lib


def _fold(inputs: Sequence[torch.Tensor], op: int, wire_dtype=None, has_scale=False, scale=1.0) -> torch.Tensor:
    ins = [t.detach().cpu().contiguous() for t in inputs]
    dt = ins[0].dtype
    assert all(t.dtype == dt and t.shape == ins[0].shape for t in ins), "ranks disagree on dtype/shape"
    out = torch.empty_like(ins[0])
    ptrs = (ctypes.c_void_p * len(ins))(*[t.data_ptr() for t in ins])
    wire = DT[dt] if wire_dtype is None else DT[wire_dtype]
    rc = lib().oracle_reduce(DT[dt], wire, op, len(ins), ins[0].numel(), ptrs, out.data_ptr(), int(has_scale), float(scale))
    if rc != 0:
        raise ValueError(f"oracle does not support dtype={dt} wire={wire_dtype} op={op}")
    return out


def allreduce(inputs: Sequence[torch.Tensor], op: int = SUM) -> torch.Tensor:
    """Value every rank holds after the allreduce."""
    return _fold(inputs, op)


def reduce(inputs: Sequence[torch.Tensor], op: int = SUM) -> torch.Tensor:
    """Value the root holds after reduce (other ranks keep their input)."""
    return _fold(inputs, op)


def allreduce_scaled(inputs: Sequence[torch.Tensor], wire_dtype=None, scale: float = 1.0) -> torch.Tensor:
    """Fused gradient reduction: SUM in fp32 over wire-rounded contributions, times scale, rounded to
    the wire type, stored in the bucket dtype."""
    return _fold(inputs, SUM, wire_dtype=wire_dtype, has_scale=True, scale=scale)


def broadcast(inputs: Sequence[torch.Tensor], root: int) -> torch.Tensor:
    return inputs[root].detach().cpu().clone()


def allgather(inputs: Sequence[torch.Tensor]) -> List[torch.Tensor]:
    return [t.detach().cpu().clone() for t in inputs]


def reducescatter(input_lists: Sequence[Sequence[torch.Tensor]], op: int = SUM) -> List[torch.Tensor]:
    """input_lists[r][j] = rank r's contribution to rank j.  Returns out[j] for every rank j."""
    world = len(input_lists)
    return [_fold([input_lists[r][j] for r in range(world)], op) for j in range(world)]
