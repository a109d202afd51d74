"""Ray Train integration (R3): a TorchConfig-shaped backend config and prepare_model().

Mirrors python/ray/train/torch/config.py (`TorchConfig` :41-70, `_setup_torch_process_group`
:73-128, `_TorchBackend.on_start` :163-212) and python/ray/train/v2/torch/train_loop_utils.py
(`prepare_model` :166-248).  Ray Train's plugin seam is `BackendConfig.backend_cls` ->
`Backend.on_start/on_training_start/on_shutdown` (python/ray/train/backend.py:15-59); the
classes below keep those names and signatures and only rely on the three WorkerGroup methods the
reference backend itself uses (`execute`, `execute_single`, `__len__`), so they plug into Ray
Train when Ray is installed and into any stand-in worker group otherwise (tests, bench.py).

What changes against the reference: torch.distributed is still initialised (DDP needs a process
group for its one-off parameter broadcast), but every per-step gradient reduction goes through
the fused peer-memory kernel registered as DDP's comm hook — NCCL is off the hot path.
"""
import logging
import os
import socket
from dataclasses import dataclass
from datetime import timedelta
from typing import Any, Dict, Optional, Union

import torch
import torch.distributed as dist

from . import ddp_hook

logger = logging.getLogger(__name__)


@dataclass
class B200TorchConfig:
    """Drop-in for ray.train.torch.TorchConfig.  `backend` is the c10d backend used for the
    control-plane process group (nccl when GPUs are present); `grad_wire` selects what crosses
    NVLink in the fused gradient reduction: "fp32" (the default: exact torch-DDP default-reducer
    semantics, so switching the import does not change training numerics), or — opt-in, like
    registering torch's bf16_compress_hook — "bf16" / "fp16" (16-bit wire, fp32 accumulate)."""

    backend: Optional[str] = None
    init_method: str = "env"
    timeout_s: int = 1800
    grad_wire: str = "fp32"

    @property
    def backend_cls(self):
        return _B200TorchBackend

    @property
    def train_func_context(self):
        return _DeviceContext


class _DeviceContext:
    def __enter__(self):
        if torch.cuda.is_available():
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))

    def __exit__(self, *exc):
        return False


def _free_address():
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return "127.0.0.1", s.getsockname()[1]


def _setup_torch_process_group(backend: str, world_rank: int, world_size: int, init_method: str, timeout_s: int = 1800,
                               grad_wire: str = "fp32"):
    """Connect torch.distributed (reference config.py:73-128) and remember the gradient wire type."""
    if backend == "nccl" and "TORCH_NCCL_ASYNC_ERROR_HANDLING" not in os.environ and "TORCH_NCCL_BLOCKING_WAIT" not in os.environ:
        os.environ["TORCH_NCCL_ASYNC_ERROR_HANDLING"] = "1"
    os.environ["B200COLL_GRAD_WIRE"] = grad_wire
    dist.init_process_group(backend=backend, init_method=init_method, rank=world_rank, world_size=world_size,
                            timeout=timedelta(seconds=timeout_s))


class _B200TorchBackend:
    share_cuda_visible_devices: bool = True  # peers must be able to map each other's HBM

    def on_start(self, worker_group, backend_config: B200TorchConfig):
        backend = backend_config.backend or ("nccl" if torch.cuda.is_available() else "gloo")
        addr, port = worker_group.execute_single(0, _free_address)
        if backend_config.init_method == "env":

            def set_env_vars(addr, port):
                os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = addr, str(port)

            worker_group.execute(set_env_vars, addr=addr, port=port)
            url = "env://"
        elif backend_config.init_method == "tcp":
            url = f"tcp://{addr}:{port}"
        else:
            raise ValueError(f"The provided init_method ({backend_config.init_method}) is not supported. "
                             "Must be either 'env' or 'tcp'.")
        n = len(worker_group)
        futures = [worker_group.execute_single_async(i, _setup_torch_process_group, backend=backend, world_rank=i,
                                                     world_size=n, init_method=url, timeout_s=backend_config.timeout_s,
                                                     grad_wire=backend_config.grad_wire) for i in range(n)]
        worker_group.wait(futures)

    def on_training_start(self, worker_group, backend_config):
        return None

    def on_shutdown(self, worker_group, backend_config):
        def _shutdown():
            if dist.is_initialized():
                dist.destroy_process_group()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()

        worker_group.execute(_shutdown)


def get_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def prepare_model(model: torch.nn.Module, move_to_device: Union[bool, torch.device] = True,
                  parallel_strategy: Optional[str] = "ddp", parallel_strategy_kwargs: Optional[Dict[str, Any]] = None,
                  grad_wire: Optional[str] = None, wrap_single: bool = False) -> torch.nn.Module:
    """ray.train.torch.prepare_model with the fused gradient reduction attached.

    Same arguments as the reference (v2/torch/train_loop_utils.py:166-248); `grad_wire` overrides
    the backend config's wire type, `wrap_single` wraps in DDP even at world size 1 (the reference
    returns the bare model there).  The returned module carries `.b200_grad_state`.
    """
    parallel_strategy_kwargs = dict(parallel_strategy_kwargs or {})
    device = move_to_device if isinstance(move_to_device, torch.device) else get_device()
    if device.type == "cuda":
        torch.cuda.set_device(device)
    if move_to_device:
        model = model.to(device)
    world_size = dist.get_world_size() if dist.is_initialized() else 1
    if parallel_strategy and (world_size > 1 or wrap_single):
        if parallel_strategy != "ddp":
            raise RuntimeError("The B200 backend accelerates the DDP gradient path; use parallel_strategy='ddp'.")
        if device.type != "cuda":
            raise RuntimeError("The B200 backend needs CUDA devices; there is no CPU fallback.")
        from torch.nn.parallel import DistributedDataParallel

        kwargs = {"device_ids": [device], "output_device": device, **parallel_strategy_kwargs}
        model = DistributedDataParallel(model, **kwargs)
        wire = grad_wire or os.environ.get("B200COLL_GRAD_WIRE", "fp32")
        if wire not in ("fp32", None):
            logger.warning("B200 gradient reduction uses a %s wire (fp32 accumulate): gradients are rounded to %s on the "
                           "way across NVLink, like torch's %s_compress_hook. Use grad_wire='fp32' for the exact "
                           "default-reducer numerics.", wire, wire, wire)
        model.b200_grad_state = ddp_hook.register(model, wire=wire)
    return model
