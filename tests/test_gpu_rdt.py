"""N1 on real GPUs (>= 2): RDT tensor transport over the B200 backend, one process per GPU."""
import pytest
import torch

from mini_actor import get, spawn
from workers import GPUWorker

from ant_ray_b200 import experimental_collective as xc
from ant_ray_b200.rdt_transport import B200TensorTransport

pytestmark = pytest.mark.gpu


def _send_gpu(self, meta):
    tensors = [torch.arange(12, dtype=torch.float32, device="cuda").reshape(3, 4), torch.full((100_000,), 3, dtype=torch.bfloat16, device="cuda")]
    B200TensorTransport().send_multiple_tensors(tensors, None, meta)
    torch.cuda.synchronize()
    return True


def _recv_gpu(self, specs, meta):
    bufs = [torch.empty(shape, dtype=dtype, device="cuda") for shape, dtype in specs]
    B200TensorTransport().recv_multiple_tensors(bufs, "obj", None, meta)
    torch.cuda.synchronize()
    return [b.cpu() for b in bufs]


def test_rdt_transfer_between_two_gpu_actors(store_dir):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    xc.set_runtime(get)
    env = {"B200COLL_STORE": f"file://{store_dir}", "B200COLL_TIMEOUT_MS": "20000", "B200COLL_RENDEZVOUS_TIMEOUT_S": "60"}
    actors = [spawn(GPUWorker, i, env=env, start_method="spawn") for i in range(2)]
    try:
        g = xc.create_collective_group(actors, "b200", name="rdt-gpu")
        tt = B200TensorTransport()
        assert tt.actor_has_tensor_transport(actors[0])
        meta = tt.get_communicator_metadata(actors[0], actors[1], "b200")
        assert (meta.src_rank, meta.dst_rank) == (0, 1)
        specs = [((3, 4), torch.float32), ((100_000,), torch.bfloat16)]
        _, got = get([actors[0].__ray_call__.remote(_send_gpu, meta), actors[1].__ray_call__.remote(_recv_gpu, specs, meta)])
        assert torch.equal(got[0], torch.arange(12, dtype=torch.float32).reshape(3, 4))
        assert torch.equal(got[1], torch.full((100_000,), 3, dtype=torch.bfloat16))
        xc.destroy_collective_group(g)
    finally:
        for h in xc.get_collective_groups([]):
            xc.RemoteCommunicatorManager.get().remove(h.name)
        for a in actors:
            a.kill()
