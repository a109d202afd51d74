"""Communicator lifecycle for compiled graphs: per-process context + driver-side orchestration.

Restates `ChannelContext` (python/ray/experimental/channel/common.py:120-174: the per-process
`communicators` map and the driver's `communicator_handles`), and `_init_communicator`,
`_do_init_communicator`, `_do_destroy_communicator`, `_destroy_communicator`
(torch_tensor_accelerator_channel.py:652-690, 738-872):

  driver:  group_id = init_communicator(actors)             # B200Communicator on every actor
           group_id = init_communicator(actors, custom)     # a user Communicator instance, pickled to
                                                            # every actor, then .initialize(rank)
           destroy_communicator(group_id)
  actor:   ChannelContext.get_current().communicators[group_id]

Ranks are positions in `actors` (or the custom communicator's own `get_rank`).  A communicator id is
generated on the first actor, as the reference does, so the driver need not share a node with
the group.  Code runs on actors through `__ray_call__`; results are resolved with `ray.get` or with
the resolver installed by `experimental_collective.set_runtime`.
"""
import logging
import threading
import uuid
from typing import Dict, Optional

from . import experimental_collective as _xc

logger = logging.getLogger(__name__)


class ChannelContext:
    _current: Optional["ChannelContext"] = None
    _lock = threading.Lock()

    def __init__(self):
        self.communicators: Dict[str, object] = {}          # inside an actor: group id -> Communicator
        self.communicator_handles: Dict[str, list] = {}     # on the driver: group id -> actor handles

    @classmethod
    def get_current(cls) -> "ChannelContext":
        with cls._lock:
            if cls._current is None:
                cls._current = cls()
            return cls._current


def _do_get_unique_communication_id(self, communicator_cls=None):
    if communicator_cls is None:  # the registered (or default) accelerator context decides, as in the reference
        from .accelerator_context import AcceleratorContext

        return AcceleratorContext.get().generate_communicator_id()
    return communicator_cls.generate_communicator_id()


def _do_init_communicator(self, group_id, world_size, comm_id, rank, actor_handles, use_communication_streams,
                          custom_communicator=None, communicator_cls=None):
    """Runs on every actor (reference torch_tensor_accelerator_channel.py:652-680).  The default path builds
    the class the accelerator-context registry names — B200Communicator once `register_b200()` has run, or by
    default on a CUDA device outside Ray — exactly the way the reference calls it: positional
    (world_size, comm_id, rank, actor_handles, current stream, use_communication_streams).
    `communicator_cls` overrides the registry for one group (tests)."""
    from .accelerator_context import AcceleratorContext

    ctx = ChannelContext.get_current()
    if custom_communicator is not None:
        custom_communicator.initialize(rank)
        ctx.communicators[group_id] = custom_communicator
        return rank
    actx = AcceleratorContext.get()
    assert actx.accelerator_count > 0, "Actors participating in Communication group must have at least one Accelerator assigned"
    if communicator_cls is not None:
        ctx.communicators[group_id] = communicator_cls(world_size, comm_id, rank, actor_handles, actx.current_stream(),
                                                       use_communication_streams)
    else:
        ctx.communicators[group_id] = actx.create_communicator(world_size, comm_id, rank, actor_handles, actx.current_stream(),
                                                               use_communication_streams)
    return rank


def _do_destroy_communicator(self, group_id):
    ctx = ChannelContext.get_current()
    if group_id in ctx.communicators:
        # the entry stays in the map: a task loop may still hold it and must see RayChannelError
        ctx.communicators[group_id].destroy()
    return True


def init_communicator(actors: list, custom_communicator=None, use_communication_streams: bool = False,
                      communicator_cls=None) -> str:
    """Create (or adopt) a communicator on every actor; returns the group id."""
    keys = {_xc._actor_key(a) for a in actors}
    assert len(keys) == len(actors), "Actors must be unique"
    comm_id = _xc._resolve(actors[0].__ray_call__.remote(_do_get_unique_communication_id,
                                                        type(custom_communicator) if custom_communicator is not None else communicator_cls))
    group_id = uuid.uuid4().hex
    world_size = len(actors)
    if custom_communicator is not None:
        ranks = [custom_communicator.get_rank(a) for a in actors]
        assert sorted(ranks) == list(range(world_size)), "custom communicator must rank every actor exactly once"
    else:
        ranks = list(range(world_size))
    _xc._resolve([a.__ray_call__.remote(_do_init_communicator, group_id, world_size, comm_id, r, actors, use_communication_streams,
                                        custom_communicator, communicator_cls) for r, a in zip(ranks, actors)])
    ChannelContext.get_current().communicator_handles[group_id] = (
        list(custom_communicator.get_actor_handles()) if custom_communicator is not None else list(actors))
    return group_id


def destroy_communicator(group_id: str) -> None:
    ctx = ChannelContext.get_current()
    actors = ctx.communicator_handles.pop(group_id, None)
    if actors is None:
        return
    _xc._resolve([a.__ray_call__.remote(_do_destroy_communicator, group_id) for a in actors])
