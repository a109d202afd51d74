"""Driver-side collective-group registry (SURVEY.md section 8f, row N1 prerequisite).

Restates python/ray/experimental/collective/collective.py:22-220 and communicator.py:23-64:
`create_collective_group(actors, backend, name)` initialises the group on every actor (rank =
position in the list) and remembers a `CommunicatorHandle`; `get_collective_groups(actors,
backend)` finds the groups a set of actors belongs to; `destroy_collective_group` tears one down.
Compiled graphs and RDT (GPU objects) look groups up through this registry.

Running code on an actor uses the actor's `__ray_call__(fn, *args)` hook, as the reference does;
`ray.get` resolves the results inside Ray, and `set_runtime(get_fn)` installs another resolver for
stand-in actor runtimes (the test-suite's mini actors).
"""
import threading
import uuid
from typing import Callable, Dict, List, Optional, Union

from . import collective as _col
from .types import Backend

_get: Optional[Callable] = None


def set_runtime(get_fn: Callable) -> None:
    """Install the function that resolves the futures returned by `actor.__ray_call__.remote`."""
    global _get
    _get = get_fn


def _resolve(refs):
    if _get is not None:
        return _get(refs)
    import ray

    return ray.get(refs)


def _actor_key(actor):
    return getattr(actor, "_ray_actor_id", id(actor))


class CommunicatorHandle:
    """What the driver remembers about a group: its actors (rank order), name and backend."""

    def __init__(self, actors: list, name: str, backend: str):
        self._actors, self._name, self._backend = list(actors), name, Backend(backend)

    def get_rank(self, actor) -> int:
        for i, a in enumerate(self._actors):
            if a is actor or a == actor:
                return i
        return -1

    actors = property(lambda self: self._actors[:])
    name = property(lambda self: self._name)
    backend = property(lambda self: self._backend)


class RemoteCommunicatorManager:
    _instance = None
    _lock = threading.Lock()

    def __init__(self):
        self._groups: Dict[str, CommunicatorHandle] = {}

    @classmethod
    def get(cls) -> "RemoteCommunicatorManager":
        with cls._lock:
            if cls._instance is None:
                cls._instance = cls()
            return cls._instance

    def add(self, handle: CommunicatorHandle):
        self._groups[handle.name] = handle

    def remove(self, name: str) -> Optional[CommunicatorHandle]:
        return self._groups.pop(name, None)

    def find(self, actors=None, backend=None) -> List[CommunicatorHandle]:
        wanted = {_actor_key(a) for a in (actors or [])}
        return [g for g in self._groups.values()
                if wanted <= {_actor_key(a) for a in g.actors} and (backend is None or g.backend == backend)]


def get_collective_groups(actors: list, backend: Optional[str] = None) -> List[CommunicatorHandle]:
    """Groups that contain every one of `actors` (optionally only those of `backend`)."""
    return RemoteCommunicatorManager.get().find(actors, Backend(backend) if backend is not None else None)


def create_collective_group(actors: list, backend: str, name: Optional[str] = None) -> CommunicatorHandle:
    """Initialise a collective group on `actors`; returns once every actor has joined."""
    manager = RemoteCommunicatorManager.get()
    name = name or uuid.uuid4().hex
    backend = Backend(backend)
    keys = [_actor_key(a) for a in actors]
    if len(set(keys)) != len(keys):
        raise ValueError(f"All actors must be unique, got: {actors}")
    for actor in actors:
        if manager.find([actor], backend):
            raise RuntimeError(f"Actor {actor} already in group for backend {backend}. Actors can currently only "
                               "participate in at most one group per backend.")
    world_size = len(actors)

    def _join(self, rank: int):
        _col.init_collective_group(world_size, rank, backend, group_name=name)
        return rank

    _resolve([actor.__ray_call__.remote(_join, rank) for rank, actor in enumerate(actors)])
    handle = CommunicatorHandle(actors, name, backend)
    manager.add(handle)
    return handle


def destroy_collective_group(group_or_name: Union[CommunicatorHandle, str]) -> None:
    if isinstance(group_or_name, CommunicatorHandle):
        name = group_or_name.name
    elif isinstance(group_or_name, str):
        name = group_or_name
    else:
        raise ValueError("Expected CommunicatorHandle or str (group name).")
    group = RemoteCommunicatorManager.get().remove(name)
    if group is None:
        raise ValueError(f"No group with name {name} found.")

    def _leave(self):
        _col.destroy_collective_group(name)

    _resolve([actor.__ray_call__.remote(_leave) for actor in group.actors])


def destroy_all_collective_groups() -> None:
    for g in RemoteCommunicatorManager.get().find():
        destroy_collective_group(g.name)
