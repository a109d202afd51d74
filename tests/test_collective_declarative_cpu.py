"""The declarative group path (`create_collective_group` on the driver, lazy creation from the `Info` actor in
`get_group_handle` on each member — reference python/ray/util/collective/collective.py:211-290, 741-789 and
util.py:54-84) executed against a minimal in-process stand-in for the `ray` module: Ray itself is not installed
in this image, so the control flow around `ray.get_actor` / `ray.get` / the worker's actor id is what is pinned
here (the arithmetic behind a group is covered by the other suites).
"""
import sys
import types as pytypes

import pytest


class _Call:
    def __init__(self, fn):
        self._fn = fn

    def remote(self, *a, **kw):
        return self._fn(*a, **kw)   # an "object ref" is the value itself


class _Handle:
    def __init__(self, obj):
        self._obj = obj

    def __getattr__(self, name):
        return _Call(getattr(self._obj, name))


def make_fake_ray():
    ray = pytypes.ModuleType("ray")
    named = {}
    state = pytypes.SimpleNamespace(actor_id=None)
    ray.WORKER_MODE, ray.SCRIPT_MODE = 1, 0

    class _Options:
        def __init__(self, cls, name=None):
            self.cls, self.name = cls, name

        def remote(self, *a, **kw):
            h = _Handle(self.cls(*a, **kw))
            if self.name is not None:
                named[self.name] = h
            return h

    class _ActorClass:
        def __init__(self, cls):
            self.cls = cls

        def options(self, name=None, lifetime=None, **_kw):
            return _Options(self.cls, name)

        def remote(self, *a, **kw):
            return _Options(self.cls).remote(*a, **kw)

    def get_actor(name):
        if name not in named:
            raise ValueError(f"Failed to look up actor with name '{name}'")
        return named[name]

    def kill(handle):
        for k in [k for k, v in named.items() if v is handle]:
            del named[k]

    worker = pytypes.SimpleNamespace(mode=ray.SCRIPT_MODE,
                                     core_worker=pytypes.SimpleNamespace(get_actor_id=lambda: state.actor_id))
    ray.remote = lambda cls: _ActorClass(cls)
    ray.get = lambda x: x
    ray.get_actor = get_actor
    ray.kill = kill
    ray.is_initialized = lambda: True
    ray._private = pytypes.SimpleNamespace(worker=pytypes.SimpleNamespace(global_worker=worker))
    ray._test = pytypes.SimpleNamespace(named=named, state=state, worker=worker)
    return ray


class _Group:
    destroyed = 0

    def __init__(self, world_size, rank, group_name, gloo_timeout):
        self.world_size, self.rank, self.group_name, self.gloo_timeout = world_size, rank, group_name, gloo_timeout

    def destroy_group(self):
        _Group.destroyed += 1


@pytest.fixture
def col(monkeypatch):
    fake = make_fake_ray()
    monkeypatch.setitem(sys.modules, "ray", fake)
    monkeypatch.delitem(sys.modules, "ant_ray_b200._ray_actors", raising=False)
    import ant_ray_b200.collective as c
    from ant_ray_b200.types import Backend

    monkeypatch.setattr(c, "ray", fake)
    monkeypatch.setattr(c, "_RAY_AVAILABLE", True)
    monkeypatch.setitem(c._BACKENDS, Backend.GLOO, _Group)
    monkeypatch.setattr(c, "_group_mgr", c.GroupManager())
    yield c, fake
    sys.modules.pop("ant_ray_b200._ray_actors", None)


def _actors(*ids):
    return [pytypes.SimpleNamespace(_ray_actor_id=i) for i in ids]


def test_member_builds_its_group_from_the_info_actor(col):
    c, ray = col
    c.create_collective_group(_actors("id-a", "id-b", "id-c"), 3, [2, 0, 1], backend="gloo", group_name="decl", gloo_timeout=1234)
    assert "info_decl" in ray._test.named
    # a second declaration of the same group is refused (reference :229-234)
    with pytest.raises(RuntimeError, match="twice"):
        c.create_collective_group(_actors("id-a", "id-b", "id-c"), 3, [2, 0, 1], backend="gloo", group_name="decl")
    # the driver itself may not call collectives
    with pytest.raises(RuntimeError, match="inside a Ray actor"):
        c.get_group_handle("decl")
    # member with actor id "id-b" is rank 0, "id-a" rank 2
    ray._test.worker.mode = ray.WORKER_MODE
    ray._test.state.actor_id = "id-b"
    g = c.get_group_handle("decl")
    assert (g.world_size, g.rank, g.group_name, g.gloo_timeout) == (3, 0, "decl", 1234)
    assert c.get_group_handle("decl") is g and c.is_group_initialized("decl")
    assert c.get_rank("decl") == 0 and c.get_collective_group_size("decl") == 3
    # destroying the group also removes the detached Info actor (reference :150-156)
    before = _Group.destroyed
    c.destroy_collective_group("decl")
    assert _Group.destroyed == before + 1 and "info_decl" not in ray._test.named and not c.is_group_initialized("decl")


def test_unknown_group_falls_back_to_the_environment_then_fails(col, monkeypatch):
    c, ray = col
    ray._test.worker.mode = ray.WORKER_MODE
    ray._test.state.actor_id = "id-x"
    with pytest.raises(RuntimeError, match="not initialized"):
        c.get_group_handle("nobody")
    for k, v in {"collective_group_name": "envg", "collective_rank": "1", "collective_world_size": "2",
                 "collective_backend": "gloo", "collective_gloo_timeout": "777"}.items():
        monkeypatch.setenv(k, v)
    g = c.get_group_handle("envg")
    assert (g.world_size, g.rank, g.gloo_timeout) == (2, 1, 777)
    c.destroy_collective_group("envg")


@pytest.mark.parametrize("world_size,ranks,n_actors,msg", [
    (2, [0], 2, "correspond to one rank"),
    (2, [0, 0], 2, "permutation"),
    (0, [], 0, "greater than zero"),
    (1, [0, 1], 2, "greater than world_size"),
])
def test_declaration_is_validated_like_the_reference(col, world_size, ranks, n_actors, msg):
    c, _ray = col
    with pytest.raises(RuntimeError, match=msg):
        c.create_collective_group(_actors(*[f"id-{i}" for i in range(n_actors)]), world_size, ranks, backend="gloo", group_name="bad")
