"""N1: RDT tensor transport + driver-side group registry, world_size-2 over gloo on CPU.

Ported shape of the reference's RDT collective-transport flow
(gpu_object_manager/collective_tensor_transport.py: metadata -> communicator lookup -> per-tensor
send/recv) and of python/ray/tests/test_experimental_collective.py's group lifecycle checks
(create twice -> error, destroy -> actors reusable)."""
import pytest
import torch

from mini_actor import get, spawn
from workers import Worker

from ant_ray_b200 import experimental_collective as xc
from ant_ray_b200.rdt_transport import B200TensorTransport, TensorTransportMetadata


@pytest.fixture
def actors(store_dir):
    xc.set_runtime(get)
    env = {"B200COLL_STORE": f"file://{store_dir}"}
    made = [spawn(Worker, env=env) for _ in range(3)]
    yield made
    for g in xc.get_collective_groups([]):
        xc.RemoteCommunicatorManager.get().remove(g.name)
    for a in made:
        a.kill()


def _send(self, tensors, meta):
    B200TensorTransport("gloo").send_multiple_tensors(tensors, None, meta)
    return True


def _recv(self, specs, meta):
    bufs = [torch.empty(shape, dtype=dtype) for shape, dtype in specs]
    B200TensorTransport("gloo").recv_multiple_tensors(bufs, "obj-1", None, meta)
    return bufs


def test_group_registry_and_transfer(actors):
    a0, a1, a2 = actors
    tt = B200TensorTransport("gloo")
    assert not tt.actor_has_tensor_transport(a0)
    with pytest.raises(ValueError):
        tt.get_communicator_metadata(a0, a1, "gloo")  # no communicator yet
    g = xc.create_collective_group([a0, a1], "gloo", name="rdt")
    assert tt.actor_has_tensor_transport(a0) and not tt.actor_has_tensor_transport(a2)
    assert (g.get_rank(a0), g.get_rank(a1), g.get_rank(a2)) == (0, 1, -1)
    with pytest.raises(RuntimeError):
        xc.create_collective_group([a1, a2], "gloo")  # a1 is already in a gloo group
    with pytest.raises(ValueError):
        xc.create_collective_group([a2, a2], "gloo")
    meta = tt.get_communicator_metadata(a1, a0, "gloo")
    assert (meta.communicator_name, meta.src_rank, meta.dst_rank) == ("rdt", 1, 0)

    payload = [torch.arange(12, dtype=torch.float32).reshape(3, 4), torch.ones(5, dtype=torch.int64) * 7]
    tmeta = tt.extract_tensor_transport_metadata("obj-1", payload)
    assert isinstance(tmeta, TensorTransportMetadata) and tmeta.tensor_device.type == "cpu"
    assert [tuple(s) for s, _ in tmeta.tensor_meta] == [(3, 4), (5,)]
    refs = [a1.__ray_call__.remote(_send, payload, meta), a0.__ray_call__.remote(_recv, [(tuple(s), d) for s, d in tmeta.tensor_meta], meta)]
    _, got = get(refs)
    assert torch.equal(got[0], payload[0]) and torch.equal(got[1], payload[1])

    xc.destroy_collective_group(g)
    assert xc.get_collective_groups([a0]) == []
    with pytest.raises(ValueError):
        xc.destroy_collective_group("rdt")
    # actors are reusable after destroy (same pair: the gloo backend keeps one default process group
    # per process, exactly like the reference's TorchGLOOGroup)
    g2 = xc.create_collective_group([a0, a1], "gloo", name="rdt2")
    assert g2.get_rank(a1) == 1
    xc.destroy_all_collective_groups()
    assert xc.get_collective_groups([]) == []


def test_transport_contract():
    tt = B200TensorTransport()
    assert tt.tensor_transport_backend == "B200"
    assert B200TensorTransport.is_one_sided() is False and B200TensorTransport.can_abort_transport() is False
    assert tt.extract_tensor_transport_metadata("x", []).tensor_meta == []
    tt.garbage_collect("x", TensorTransportMetadata())
