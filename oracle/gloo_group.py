"""Restatement of the reference's CPU backend, TorchGLOOGroup, over the real gloo library.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle_reduce.c).  It follows
python/ray/util/collective/collective_group/torch_gloo_collective_group.py op for op:
  __init__       :40-94   (default process group once per process, one subgroup per extra group)
  allreduce      :147-156
  reduce         :160-179 (non-root ranks reduce into a clone so their tensor is untouched)
  allgather      :181-191
  broadcast      :193-197
  reducescatter  :199-221 (gloo has no reduce_scatter: W allreduces, then a local copy)
  send / recv    :223-229
but takes its rendezvous address from a `Store` of ant_ray_b200.rendezvous instead of Ray's
internal KV (collective.py:93-110), so that it runs without Ray.  `register()` plugs it into
ant_ray_b200.collective's backend registry under "GLOO" — that is how configs[0]
("ray.util.collective.allreduce world_size=2 gloo backend on CPU") is exercised on CPU.
"""
import os
import socket
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ant_ray_b200 import rendezvous
from ant_ray_b200.types import Backend, ReduceOp

_TORCH_OP = {ReduceOp.SUM: dist.ReduceOp.SUM, ReduceOp.PRODUCT: dist.ReduceOp.PRODUCT,
             ReduceOp.MIN: dist.ReduceOp.MIN, ReduceOp.MAX: dist.ReduceOp.MAX}


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _as_tensor(x) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        return torch.from_numpy(x)  # zero-copy: the result lands in the caller's ndarray
    raise ValueError(f"torch_gloo group only accepts torch.Tensor or numpy.ndarray, received {type(x)}")


class GlooGroup:
    def __init__(self, world_size: int, rank: int, group_name: str, gloo_timeout: Optional[int] = None,
                 store: Optional[rendezvous.Store] = None):
        self._world_size, self._rank, self._group_name = world_size, rank, group_name
        timeout_s = (gloo_timeout or 30000) / 1000.0
        if not dist.is_initialized():
            store = store if store is not None else rendezvous.default_store()
            # the n-th incarnation of a group name gets its own key, otherwise a re-created group could
            # read the previous incarnation's (dead) address
            from ant_ray_b200.b200_group import next_comm_key

            key = "collective_group_master_address_" + next_comm_key(group_name)
            if rank == 0:
                store.set(key, f"127.0.0.1:{_free_port()}".encode())
            addr, port = store.get(key, timeout_s).decode().split(":")
            os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = addr, port
            dist.init_process_group(backend="gloo", init_method="env://", world_size=world_size, rank=rank)
        self._is_default = group_name == "default"
        self._pg = dist.group.WORLD if self._is_default else dist.new_group(ranks=list(range(world_size)), backend="gloo")
        self._timeout_ms = gloo_timeout if gloo_timeout is not None else 30000

    rank = property(lambda self: self._rank)
    world_size = property(lambda self: self._world_size)
    group_name = property(lambda self: self._group_name)

    @classmethod
    def backend(cls):
        return Backend.GLOO

    def destroy_group(self):
        if self._is_default:
            dist.destroy_process_group()
        elif self._pg is not None:
            dist.destroy_process_group(self._pg)

    @staticmethod
    def _one(tensors) -> torch.Tensor:
        assert isinstance(tensors, list) and len(tensors) == 1
        return _as_tensor(tensors[0])

    @staticmethod
    def _many(tensor_lists) -> List[torch.Tensor]:
        assert isinstance(tensor_lists, list) and len(tensor_lists) == 1
        return [_as_tensor(t) for t in tensor_lists[0]]

    def allreduce(self, tensors, opts=None):
        op = _TORCH_OP[opts.reduceOp] if opts is not None else dist.ReduceOp.SUM
        dist.all_reduce(self._one(tensors), op=op, group=self._pg)

    def barrier(self, opts=None):
        dist.barrier(group=self._pg)

    def reduce(self, tensors, opts):
        t = self._one(tensors)
        target = t if self._rank == opts.root_rank else t.detach().clone()
        dist.reduce(target, dst=opts.root_rank, op=_TORCH_OP[opts.reduceOp], group=self._pg)

    def allgather(self, tensor_lists, tensors, opts=None):
        dist.all_gather(self._many(tensor_lists), self._one(tensors), group=self._pg)

    def broadcast(self, tensors, opts):
        dist.broadcast(self._one(tensors), src=opts.root_rank, group=self._pg)

    def reducescatter(self, tensors, tensor_lists, opts):
        ins, out = self._many(tensor_lists), self._one(tensors)
        if out.shape != ins[self._rank].shape:
            raise ValueError(f"Output tensor has wrong shape {out.shape}, expected {ins[self._rank].shape}")
        for t in ins:
            dist.all_reduce(t, op=_TORCH_OP[opts.reduceOp], group=self._pg)
        if out.data_ptr() != ins[self._rank].data_ptr():
            out.copy_(ins[self._rank])

    def send(self, tensors, opts):
        dist.send(self._one(tensors), dst=opts.dst_rank)

    def recv(self, tensors, opts):
        dist.recv(self._one(tensors), src=opts.src_rank)


def register(store: Optional[rendezvous.Store] = None):
    """Make backend="gloo" available in ant_ray_b200.collective (tests and CPU baseline only)."""
    from ant_ray_b200 import collective

    collective.register_backend(Backend.GLOO, lambda w, r, name, timeout: GlooGroup(w, r, name, timeout, store))
