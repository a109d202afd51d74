"""The ray.util.collective API surface, served by the B200 peer-memory backend.

Function names, argument order, defaults and error behaviour follow the reference
python/ray/util/collective/collective.py (init_collective_group :171-208, allreduce :328-343,
barrier :368, reduce :381, broadcast :443, allgather :493, reducescatter :542, send :601,
recv :664, get_group_handle :741-789, input validation :792-874) so actor code and the
reference's tests port by changing one import.

What is different:
  * `backend` defaults to B200 (the string "nccl" is accepted as an alias);
  * group creation goes through a registry (`register_backend`) instead of an if/elif, which is
    also how the test-suite plugs in the gloo CPU oracle;
  * Ray is optional.  Inside Ray the declarative path (`create_collective_group`, `Info` actor)
    and the inside-actor check work as in the reference; without Ray the imperative
    `init_collective_group` path works in any process (rendezvous via rendezvous.default_store()).
"""
import logging
import os
import threading
from typing import Callable, Dict, List

import numpy as np

from . import types

logger = logging.getLogger(__name__)

try:  # Ray is the host runtime in production, but none of the arithmetic depends on it
    import ray

    _RAY_AVAILABLE = True
except ImportError:
    ray = None
    _RAY_AVAILABLE = False


def _make_b200_group(world_size, rank, group_name, gloo_timeout):
    from .b200_group import B200Group

    return B200Group(world_size, rank, group_name)


_BACKENDS: Dict[str, Callable] = {types.Backend.B200: _make_b200_group}


def register_backend(name: str, factory: Callable) -> None:
    """Register `factory(world_size, rank, group_name, gloo_timeout) -> group` for a backend name."""
    _BACKENDS[name] = factory


def b200_available() -> bool:
    try:
        from . import _native

        _native.load()
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def nccl_available() -> bool:
    """Kept for source compatibility: "nccl" names the B200 backend here."""
    return b200_available()


def gloo_available() -> bool:
    return types.Backend.GLOO in _BACKENDS


class GroupManager(object):
    """Per-process registry of collective groups (reference collective.py:71-156)."""

    def __init__(self):
        self._name_group_map = {}

    def create_collective_group(self, backend, world_size, rank, group_name, gloo_timeout):
        backend = types.Backend(backend)
        factory = _BACKENDS.get(backend)
        if factory is None:
            raise RuntimeError(f"Unexpected backend: {backend}")
        logger.debug("Creating %s group: '%s'...", backend, group_name)
        g = factory(world_size, rank, group_name, gloo_timeout)
        self._name_group_map[group_name] = g
        return g

    def is_group_exist(self, group_name):
        return group_name in self._name_group_map

    def get_group_by_name(self, group_name):
        if not self.is_group_exist(group_name):
            logger.warning("The group '{}' is not initialized.".format(group_name))
            return None
        return self._name_group_map[group_name]

    def destroy_collective_group(self, group_name):
        if not self.is_group_exist(group_name):
            logger.warning("The group '{}' does not exist.".format(group_name))
            return
        g = self._name_group_map.pop(group_name)
        g.destroy_group()
        if _RAY_AVAILABLE and ray.is_initialized():
            try:
                ray.kill(ray.get_actor("info_" + group_name))
            except ValueError:
                pass


_group_mgr = GroupManager()
_group_mgr_lock = threading.Lock()


def is_group_initialized(group_name):
    with _group_mgr_lock:
        return _group_mgr.is_group_exist(group_name)


def init_collective_group(world_size: int, rank: int, backend=types.Backend.B200, group_name: str = "default",
                          gloo_timeout: int = 30000):
    """Initialize a collective group inside an actor (or any) process."""
    _check_inside_actor()
    backend = types.Backend(backend)
    _check_backend_availability(backend)
    if not group_name:
        raise ValueError("group_name '{}' needs to be a string.".format(group_name))
    with _group_mgr_lock:
        if _group_mgr.is_group_exist(group_name):
            raise RuntimeError("Trying to initialize a group twice.")
        assert world_size > 0
        assert rank >= 0
        assert rank < world_size
        _group_mgr.create_collective_group(backend, world_size, rank, group_name, gloo_timeout)


def create_collective_group(actors, world_size: int, ranks: List[int], backend=types.Backend.B200,
                            group_name: str = "default", gloo_timeout: int = 30000):
    """Declare a list of Ray actors as a collective group (driver side; needs Ray)."""
    if not _RAY_AVAILABLE:
        raise RuntimeError("create_collective_group() needs Ray; use init_collective_group() inside each process.")
    backend = types.Backend(backend)
    _check_backend_availability(backend)
    name = "info_" + group_name
    try:
        ray.get_actor(name)
        raise RuntimeError("Trying to initialize a group twice.")
    except ValueError:
        pass
    if len(ranks) != len(actors):
        raise RuntimeError("Each actor should correspond to one rank. Got '{}' ranks but '{}' actors".format(
            len(ranks), len(actors)))
    if set(ranks) != set(range(len(ranks))):
        raise RuntimeError("Ranks must be a permutation from 0 to '{}'. Got '{}'.".format(
            len(ranks), "".join([str(r) for r in ranks])))
    if world_size <= 0:
        raise RuntimeError("World size must be greater than zero. Got '{}'.".format(world_size))
    if min(ranks) < 0:
        raise RuntimeError("Ranks must be non-negative.")
    if max(ranks) >= world_size:
        raise RuntimeError("Ranks cannot be greater than world_size.")
    from ._ray_actors import Info

    actors_id = [a._ray_actor_id for a in actors]
    info = Info.options(name=name, lifetime="detached").remote()
    ray.get([info.set_info.remote(actors_id, world_size, ranks, backend, gloo_timeout)])


def destroy_collective_group(group_name: str = "default") -> None:
    _check_inside_actor()
    with _group_mgr_lock:
        _group_mgr.destroy_collective_group(group_name)


def get_rank(group_name: str = "default") -> int:
    """Rank of this process in the group, -1 if the group does not exist here."""
    _check_inside_actor()
    with _group_mgr_lock:
        if not _group_mgr.is_group_exist(group_name):
            return -1
        return _group_mgr.get_group_by_name(group_name).rank


def get_collective_group_size(group_name: str = "default") -> int:
    """World size of the group, -1 if the group does not exist here."""
    _check_inside_actor()
    with _group_mgr_lock:
        if not _group_mgr.is_group_exist(group_name):
            return -1
        return _group_mgr.get_group_by_name(group_name).world_size


def _options(cls, **fields):
    opts = cls()
    for name, value in fields.items():
        setattr(opts, name, value)
    return opts


def _peer_rank(g, rank: int, what: str) -> int:
    """Validate a peer rank for p2p (ValueError out of range, RuntimeError for self)."""
    _check_rank_valid(g, rank)
    if rank == g.rank:
        raise RuntimeError("The {} rank '{}' is self.".format(what, rank))
    return rank


def _full_list(g, tensor_list, op_name: str):
    _check_tensor_list_input(tensor_list)
    if len(tensor_list) != g.world_size:
        raise RuntimeError("The length of the tensor list operands to {} must be equal to world_size.".format(op_name))
    return tensor_list


def allreduce(tensor, group_name: str = "default", op=types.ReduceOp.SUM):
    """Reduce `tensor` across the group; every member ends up with the result, in place."""
    _check_single_tensor_input(tensor)
    get_group_handle(group_name).allreduce([tensor], _options(types.AllReduceOptions, reduceOp=op))


def barrier(group_name: str = "default"):
    """Block until every member of the group has reached the barrier."""
    get_group_handle(group_name).barrier()


def reduce(tensor, dst_rank: int = 0, group_name: str = "default", op=types.ReduceOp.SUM):
    """Reduce `tensor` across the group onto `dst_rank` (other members keep their input)."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    _check_rank_valid(g, dst_rank)
    g.reduce([tensor], _options(types.ReduceOptions, reduceOp=op, root_rank=dst_rank, root_tensor=0))


def broadcast(tensor, src_rank: int = 0, group_name: str = "default"):
    """Copy `src_rank`'s tensor into every member's tensor."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    _check_rank_valid(g, src_rank)
    g.broadcast([tensor], _options(types.BroadcastOptions, root_rank=src_rank, root_tensor=0))


def allgather(tensor_list: list, tensor, group_name: str = "default"):
    """Gather every member's `tensor` into `tensor_list` (one slot per rank) on every member."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    g.allgather([_full_list(g, tensor_list, "allgather")], [tensor], types.AllGatherOptions())


def reducescatter(tensor, tensor_list: list, group_name: str = "default", op=types.ReduceOp.SUM):
    """Reduce the members' `tensor_list`s slot by slot; member r receives slot r in `tensor`."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    g.reducescatter([tensor], [_full_list(g, tensor_list, "reducescatter")], _options(types.ReduceScatterOptions, reduceOp=op))


def send(tensor, dst_rank: int, group_name: str = "default"):
    """Send `tensor` to `dst_rank` (pairs with a recv on that member)."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    g.send([tensor], _options(types.SendOptions, dst_rank=_peer_rank(g, dst_rank, "destination")))


def recv(tensor, src_rank: int, group_name: str = "default"):
    """Receive into `tensor` from `src_rank`."""
    _check_single_tensor_input(tensor)
    g = get_group_handle(group_name)
    g.recv([tensor], _options(types.RecvOptions, src_rank=_peer_rank(g, src_rank, "source")))


def _multigpu_unsupported(*_args, **_kwargs):
    """The reference's *_multigpu calls drive several GPUs from one process.  This backend is one
    process per GPU by design."""
    raise RuntimeError("Multigpu calls are not supported by the B200 backend: run one process (actor) per GPU.")


allreduce_multigpu = reduce_multigpu = broadcast_multigpu = allgather_multigpu = _multigpu_unsupported
reducescatter_multigpu = send_multigpu = recv_multigpu = _multigpu_unsupported


def synchronize(gpu_id: int):
    """Block the host until device `gpu_id` is idle (reference collective.py:724-738 uses cupy)."""
    import torch

    torch.cuda.synchronize(gpu_id)


def get_group_handle(group_name: str = "default"):
    """Return the group, creating it lazily from the `Info` actor (declarative path) or from the
    collective_* environment variables (reference collective.py:741-789)."""
    _check_inside_actor()
    with _group_mgr_lock:
        if not _group_mgr.is_group_exist(group_name):
            created = False
            if _RAY_AVAILABLE and ray.is_initialized():
                try:
                    mgr = ray.get_actor(name="info_" + group_name)
                    ids, world_size, rank, backend, gloo_timeout = ray.get(mgr.get_info.remote())
                    worker = ray._private.worker.global_worker
                    id_ = worker.core_worker.get_actor_id()
                    r = rank[ids.index(id_)]
                    _group_mgr.create_collective_group(backend, world_size, r, group_name, gloo_timeout)
                    created = True
                except ValueError:
                    pass
            if not created:
                if os.environ.get("collective_group_name") == group_name:
                    rank = int(os.environ["collective_rank"])
                    world_size = int(os.environ["collective_world_size"])
                    backend = os.environ["collective_backend"]
                    gloo_timeout = int(os.getenv("collective_gloo_timeout", 30000))
                    _group_mgr.create_collective_group(backend, world_size, rank, group_name, gloo_timeout)
                else:
                    raise RuntimeError("The collective group '{}' is not initialized in the process.".format(group_name))
        return _group_mgr.get_group_by_name(group_name)


def _check_single_tensor_input(tensor):
    if isinstance(tensor, np.ndarray):
        return
    if types.cupy_available() and isinstance(tensor, types.cp.ndarray):
        return
    if types.torch_available() and isinstance(tensor, types.th.Tensor):
        return
    if hasattr(tensor, "__cuda_array_interface__"):
        return
    raise RuntimeError("Unrecognized tensor type '{}'. Supported types are: np.ndarray, torch.Tensor, "
                       "cupy.ndarray.".format(type(tensor)))


def _check_backend_availability(backend):
    if backend == types.Backend.GLOO:
        if not gloo_available():
            raise RuntimeError("GLOO backend is not registered in this process.")
    elif backend == types.Backend.B200:
        if not b200_available():
            raise RuntimeError("B200 backend is not available (needs libb200coll.so and a CUDA device).")


def _check_inside_actor():
    """Inside Ray the collective API may only be used from a worker (reference :819-828).
    Without Ray every process is its own 'actor'."""
    if not _RAY_AVAILABLE or not ray.is_initialized():
        return
    worker = ray._private.worker.global_worker
    if worker.mode == ray.WORKER_MODE:
        return
    raise RuntimeError("The collective APIs shall be only used inside a Ray actor or task.")


def _check_rank_valid(g, rank: int):
    if rank < 0:
        raise ValueError("rank '{}' is negative.".format(rank))
    if rank >= g.world_size:
        raise ValueError("rank '{}' must be less than world size '{}'".format(rank, g.world_size))


def _check_tensor_list_input(tensor_list):
    if not isinstance(tensor_list, list):
        raise RuntimeError("The input must be a list of tensors. Got '{}'.".format(type(tensor_list)))
    if not tensor_list:
        raise RuntimeError("Got an empty list of tensors.")
    for t in tensor_list:
        _check_single_tensor_input(t)
