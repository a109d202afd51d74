"""Worker actors mirroring the reference's test workers
(python/ray/util/collective/tests/cpu_util.py:13-98 and tests/util.py:14-92): same method names and
defaults, driven through tests/mini_actor.py instead of ray.remote."""
import numpy as np
import torch

import ant_ray_b200.collective as col
from ant_ray_b200.types import Backend, ReduceOp


class Worker:
    """CPU worker (gloo backend, registered from the oracle package).

    `do_<op>` methods run one collective on the worker's buffers and return the buffer(s) the op wrote;
    `report_*` methods expose the group-introspection API."""

    def __init__(self):
        from oracle import gloo_group

        gloo_group.register()
        self.buffer = None
        self.list_buffer = None

    # ---- buffers ------------------------------------------------------------------------------
    def init_tensors(self):
        self.buffer = np.ones((10,), dtype=np.float32)
        self.list_buffer = [np.ones((10,), dtype=np.float32) for _ in range(2)]
        return True

    def set_buffer(self, data):
        self.buffer = data
        return self.buffer

    def get_buffer(self):
        return self.buffer

    def set_list_buffer(self, list_of_arrays, copy=False):
        self.list_buffer = [(t.copy() if isinstance(t, np.ndarray) else t.clone().detach()) for t in list_of_arrays] if copy else list_of_arrays
        return self.list_buffer

    # ---- group lifecycle ----------------------------------------------------------------------
    def init_group(self, world_size, rank, backend=Backend.B200, group_name="default"):
        col.init_collective_group(world_size, rank, backend, group_name)
        return True

    def destroy_group(self, group_name="default"):
        col.destroy_collective_group(group_name)
        return True

    # ---- collectives --------------------------------------------------------------------------
    def _run(self, fn, *args, out="buffer"):
        fn(*args)
        return [self._ret(t) for t in self.list_buffer] if out == "list" else self._ret(self.buffer)

    def do_allreduce(self, group_name="default", op=ReduceOp.SUM):
        return self._run(col.allreduce, self.buffer, group_name, op)

    def do_reduce(self, group_name="default", dst_rank=0, op=ReduceOp.SUM):
        return self._run(col.reduce, self.buffer, dst_rank, group_name, op)

    def do_broadcast(self, group_name="default", src_rank=0):
        return self._run(col.broadcast, self.buffer, src_rank, group_name)

    def do_allgather(self, group_name="default"):
        return self._run(col.allgather, self.list_buffer, self.buffer, group_name, out="list")

    def do_reducescatter(self, group_name="default", op=ReduceOp.SUM):
        return self._run(col.reducescatter, self.buffer, self.list_buffer, group_name, op)

    def do_send(self, group_name="default", dst_rank=0):
        return self._run(col.send, self.buffer, dst_rank, group_name)

    def do_recv(self, group_name="default", src_rank=0):
        return self._run(col.recv, self.buffer, src_rank, group_name)

    def do_barrier(self, group_name="default"):
        col.barrier(group_name)
        return True

    # ---- introspection ------------------------------------------------------------------------
    def report_rank(self, group_name="default"):
        return col.get_rank(group_name)

    def report_world_size(self, group_name="default"):
        return col.get_collective_group_size(group_name)

    def report_nccl_availability(self):
        return col.nccl_available()

    def report_gloo_availability(self):
        return col.gloo_available()

    def report_is_group_initialized(self, group_name="default"):
        return col.is_group_initialized(group_name)

    @staticmethod
    def _ret(t):
        return t


class GPUWorker(Worker):
    """One process per GPU; buffers are torch CUDA tensors (the reference's GPU worker uses cupy,
    tests/util.py:14-92; cupy is not installed here).  Results are returned on the host."""

    def __init__(self, device):
        super().__init__()
        torch.cuda.set_device(device)
        self.device = device

    def init_tensors(self):
        self.buffer = torch.ones((10,), dtype=torch.float32, device="cuda")
        self.list_buffer = [torch.ones((10,), dtype=torch.float32, device="cuda") for _ in range(2)]
        return True

    def set_buffer(self, data, on_gpu=True):
        self.buffer = self._to(data, on_gpu)
        return True

    def set_list_buffer(self, list_of_arrays, copy=False, on_gpu=True):
        self.list_buffer = [self._to(t, on_gpu) for t in list_of_arrays]
        return True

    def get_buffer(self):
        return self._ret(self.buffer)

    def sync(self):
        torch.cuda.synchronize()
        return True

    def _to(self, data, on_gpu):
        t = torch.from_numpy(data) if isinstance(data, np.ndarray) else data
        return t.cuda() if on_gpu else t

    @staticmethod
    def _ret(t):
        if isinstance(t, torch.Tensor) and t.is_cuda:
            torch.cuda.synchronize()
            return t.cpu()
        return t


def create_collective_workers(num_workers=2, group_name="default", backend="gloo", store_dir=None, gpu=False):
    """Spawn workers, give each a tensor set and put them in one group
    (reference cpu_util.py:101-118 / util.py:95-112)."""
    from mini_actor import get, spawn

    env = {"B200COLL_STORE": f"file://{store_dir}", "B200COLL_TIMEOUT_MS": "20000",
           "B200COLL_RENDEZVOUS_TIMEOUT_S": "60"}
    if gpu:
        actors = [spawn(GPUWorker, i, env=env, start_method="spawn") for i in range(num_workers)]
    else:
        actors = [spawn(Worker, env=env) for _ in range(num_workers)]
    get([a.init_tensors.remote() for a in actors])
    get([a.init_group.remote(num_workers, i, backend, group_name) for i, a in enumerate(actors)])
    return actors
