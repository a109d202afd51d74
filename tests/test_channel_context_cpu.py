"""Communicator lifecycle (driver-orchestrated) with stand-in actors and a recording communicator.

Follows what the reference asserts with its MockNcclGroupSet
(python/ray/experimental/collective/conftest.py:93-243): every actor gets a communicator for the
group id with its own rank, a custom communicator is adopted through initialize(rank), destroy
reaches every actor and leaves the (closed) entry in place."""
import pytest

from mini_actor import get, spawn

from ant_ray_b200 import channel_context as cc
from ant_ray_b200 import experimental_collective as xc
from ant_ray_b200.communicator import Communicator


class Recorder(Communicator):
    """Communicator double: remembers how it was built; usable as default class and as custom instance."""

    def __init__(self, world_size=None, comm_id=None, rank=None, actor_handles=None, cuda_stream=None, use_communication_streams=False):
        self.world_size, self.comm_id, self.rank, self.handles = world_size, comm_id, rank, actor_handles
        self.streams, self.closed, self.initialized_with = use_communication_streams, False, None

    def initialize(self, rank): self.initialized_with = rank
    def get_actor_handles(self): return self.handles
    def get_rank(self, actor): return [getattr(a, "_ray_actor_id", a) for a in self.handles].index(getattr(actor, "_ray_actor_id", actor))
    def get_self_rank(self): return self.rank if self.rank is not None else self.initialized_with
    def get_world_size(self): return self.world_size
    def send(self, v, p): pass
    def recv(self, s, d, p, allocator=None): pass
    recv_stream = send_stream = property(lambda self: None)
    def allgather(self, s, r): pass
    def allreduce(self, s, r, op): pass
    def reducescatter(self, s, r, op): pass
    def destroy(self): self.closed = True
    def get_transport_name(self): return "accelerator"
    @classmethod
    def generate_communicator_id(cls): return "comm-id-from-actor-0"


class Holder:
    pass


def _inspect(self, group_id):
    c = cc.ChannelContext.get_current().communicators[group_id]
    return (c.world_size, c.comm_id, c.get_self_rank(), c.streams, c.closed, len(c.handles))


_ORIGINAL_INIT = cc._do_init_communicator


def _patched_init(self, group_id, world_size, comm_id, rank, actor_handles, streams, custom, cls):
    # the default path asserts a CUDA device; on CPU build the class directly (what the assert guards)
    ctx = cc.ChannelContext.get_current()
    if custom is not None:
        return _ORIGINAL_INIT(self, group_id, world_size, comm_id, rank, actor_handles, streams, custom, cls)
    ctx.communicators[group_id] = cls(world_size, comm_id, rank, actor_handles, None, streams)
    return rank


@pytest.fixture
def actors(monkeypatch):
    xc.set_runtime(get)
    monkeypatch.setattr(cc, "_do_init_communicator", _patched_init)
    made = [spawn(Holder) for _ in range(3)]
    yield made
    for a in made:
        a.kill()


def test_default_communicator_on_every_actor(actors):
    gid = cc.init_communicator(actors, use_communication_streams=True, communicator_cls=Recorder)
    assert cc.ChannelContext.get_current().communicator_handles[gid] == actors
    infos = get([a.__ray_call__.remote(_inspect, gid) for a in actors])
    assert infos == [(3, "comm-id-from-actor-0", r, True, False, 3) for r in range(3)]
    cc.destroy_communicator(gid)
    assert gid not in cc.ChannelContext.get_current().communicator_handles
    infos = get([a.__ray_call__.remote(_inspect, gid) for a in actors])
    assert all(i[4] is True for i in infos)  # closed, but still in the map
    cc.destroy_communicator(gid)  # idempotent


def test_custom_communicator_is_adopted(actors):
    custom = Recorder(world_size=2, actor_handles=[actors[1], actors[0]])  # custom rank order: actor1 = rank 0
    gid = cc.init_communicator([actors[0], actors[1]], custom_communicator=custom)
    infos = get([a.__ray_call__.remote(_inspect, gid) for a in (actors[0], actors[1])])
    assert [i[2] for i in infos] == [1, 0]  # initialize(rank) got the communicator's own ranks
    cc.destroy_communicator(gid)


def test_duplicate_actors_rejected(actors):
    with pytest.raises(AssertionError):
        cc.init_communicator([actors[0], actors[0]], communicator_cls=Recorder)
