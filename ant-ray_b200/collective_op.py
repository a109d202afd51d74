"""Collective operation of a compiled graph: allocate the output, call the communicator.

Restates `_CollectiveOperation.execute` (python/ray/dag/collective_node.py:176-248): allgather
output `[W*n0, ...]`, reducescatter output `[n0/W, ...]` (first dimension must divide), allreduce
output like the input; several input tensors are reduced together and returned as a tuple.
The reference flattens multi-tensor inputs with `parameters_to_vector` (one extra device copy of
everything) and reduces the flat vector in place.  Here small inputs still take that single-launch
route, while large ones are reduced tensor by tensor straight into slices of one output allocation:
the communicator is out of place anyway, so the flatten copy disappears.
"""
from typing import Tuple, Union

from .types import DagReduceOp

FLATTEN_BELOW_BYTES = 1 << 20


class AllGatherOp:
    pass


class AllReduceOp:
    def __init__(self, reduceOp=DagReduceOp.SUM):
        self.reduceOp = reduceOp


class ReduceScatterOp:
    def __init__(self, reduceOp=DagReduceOp.SUM):
        self.reduceOp = reduceOp


class CollectiveOperation:
    def __init__(self, communicator, op, world_size: int):
        self._communicator, self._op, self._world_size = communicator, op, world_size

    def execute(self, *send_buf) -> Union["torch.Tensor", Tuple["torch.Tensor", ...]]:
        import torch

        if not send_buf or not all(isinstance(t, torch.Tensor) for t in send_buf):
            raise ValueError("Expected a torch tensor for each input node")
        comm, W = self._communicator, self._world_size
        if isinstance(self._op, AllGatherOp):
            assert len(send_buf) == 1
            t = send_buf[0]
            out = torch.empty((t.shape[0] * W, *t.shape[1:]), dtype=t.dtype, device=t.device)
            comm.allgather(t, out)
            return out
        if isinstance(self._op, ReduceScatterOp):
            assert len(send_buf) == 1
            t = send_buf[0]
            if t.shape[0] % W != 0:
                raise ValueError(f"Expected the first dimension of the input tensor to be divisible by the world size {W}")
            out = torch.empty((t.shape[0] // W, *t.shape[1:]), dtype=t.dtype, device=t.device)
            comm.reducescatter(t, out, self._op.reduceOp)
            return out
        if not isinstance(self._op, AllReduceOp):
            raise ValueError(f"Unknown collective operation {self._op}")
        if len(send_buf) == 1:
            out = torch.empty_like(send_buf[0])
            comm.allreduce(send_buf[0], out, self._op.reduceOp)
            return out
        if not all(t.dtype == send_buf[0].dtype for t in send_buf):
            raise ValueError(f"Expected all input tensors to have the same dtype, but got {[t.dtype for t in send_buf]}")
        total = sum(t.numel() for t in send_buf)
        flat = torch.empty(total, dtype=send_buf[0].dtype, device=send_buf[0].device)
        views, offset = [], 0
        for t in send_buf:
            views.append(flat[offset:offset + t.numel()].view(t.shape))
            offset += t.numel()
        if total * flat.element_size() < FLATTEN_BELOW_BYTES:
            for v, t in zip(views, send_buf):  # small: one launch beats several
                v.copy_(t)
            comm.allreduce(flat, flat, self._op.reduceOp)
        else:
            for v, t in zip(views, send_buf):  # large: no flatten copy, reduce straight into the slice
                comm.allreduce(t.contiguous(), v, self._op.reduceOp)
        return tuple(views)
