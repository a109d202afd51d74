"""R3 host logic on CPU: the TorchConfig-shaped backend driving a stand-in worker group (two
processes, gloo), the hook's contract with torch DDP, and the no-CPU-fallback rule.

Mirrors what the reference tests of the Train torch backend check without GPUs
(python/ray/train/tests/test_torch_trainer.py: process group comes up on every worker with the
right rank / world size, MASTER_ADDR/PORT are propagated, shutdown destroys the group).
"""
import inspect

import cloudpickle
import pytest
import torch

from mini_actor import get, spawn

from ant_ray_b200 import ddp_hook
from ant_ray_b200 import train as T


class TrainWorker:
    def run(self, payload):
        fn, args, kwargs = cloudpickle.loads(payload)
        return fn(*args, **kwargs)


class FakeWorkerGroup:
    """The three methods the reference backend uses on its worker group (execute, execute_single,
    execute_single_async) plus wait(), over tests/mini_actor.py actors."""

    def __init__(self, n):
        self.actors = [spawn(TrainWorker) for _ in range(n)]

    def __len__(self):
        return len(self.actors)

    def execute_single_async(self, i, fn, *args, **kwargs):
        return self.actors[i].run.remote(cloudpickle.dumps((fn, args, kwargs)))

    def execute_single(self, i, fn, *args, **kwargs):
        return get(self.execute_single_async(i, fn, *args, **kwargs))

    def execute(self, fn, *args, **kwargs):
        return get([self.execute_single_async(i, fn, *args, **kwargs) for i in range(len(self))])

    def wait(self, futures):
        return get(futures)

    def shutdown(self):
        for a in self.actors:
            a.kill()


def _probe():
    import os

    import torch.distributed as dist

    t = torch.ones(4) * (dist.get_rank() + 1)
    dist.all_reduce(t)
    return dist.get_rank(), dist.get_world_size(), dist.get_backend(), os.environ["MASTER_ADDR"], os.environ["B200COLL_GRAD_WIRE"], t.tolist()


def _is_init():
    import torch.distributed as dist

    return dist.is_initialized()


def test_backend_brings_up_process_group_on_every_worker():
    wg = FakeWorkerGroup(2)
    try:
        cfg = T.B200TorchConfig(backend="gloo", grad_wire="fp32", timeout_s=60)
        backend = cfg.backend_cls()
        backend.on_start(wg, cfg)
        res = wg.execute(_probe)
        assert [r[0] for r in res] == [0, 1]
        assert all(r[1] == 2 and r[2] == "gloo" and r[3] == "127.0.0.1" and r[4] == "fp32" for r in res)
        assert all(r[5] == [3.0] * 4 for r in res)
        backend.on_shutdown(wg, cfg)
        assert wg.execute(_is_init) == [False, False]
    finally:
        wg.shutdown()


def test_backend_rejects_unknown_init_method():
    wg = FakeWorkerGroup(1)
    try:
        cfg = T.B200TorchConfig(backend="gloo", init_method="carrier-pigeon")
        with pytest.raises(ValueError):
            cfg.backend_cls().on_start(wg, cfg)
    finally:
        wg.shutdown()


def test_config_defaults_match_the_reference():
    cfg = T.B200TorchConfig()
    # ray.train.torch.TorchConfig: backend=None (nccl with GPUs, gloo without), init_method="env", timeout_s=1800
    assert (cfg.backend, cfg.init_method, cfg.timeout_s) == (None, "env", 1800)
    assert cfg.grad_wire == "fp32"
    assert cfg.backend_cls is T._B200TorchBackend and T._B200TorchBackend.share_cuda_visible_devices is True


def test_hook_signature_is_what_ddp_requires():
    sig = inspect.signature(ddp_hook.b200_allreduce_hook)
    assert list(sig.parameters) == ["state", "bucket"]
    assert sig.return_annotation == torch.futures.Future[torch.Tensor]  # DDP._check_comm_hook insists on this


def test_prepare_model_has_no_cpu_path():
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    m = torch.nn.Linear(4, 4)
    # world size 1 and no wrapping requested: the model is returned as is, like the reference
    assert T.prepare_model(m, move_to_device=torch.device("cpu")) is m
    with pytest.raises(RuntimeError):
        T.prepare_model(m, move_to_device=torch.device("cpu"), wrap_single=True)
    with pytest.raises(RuntimeError):
        T.prepare_model(m, move_to_device=torch.device("cpu"), parallel_strategy="fsdp", wrap_single=True)


def test_wire_names():
    with pytest.raises(ValueError):
        ddp_hook.B200GradState(None, wire="int8")
