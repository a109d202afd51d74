"""Accelerator context: how the compiled-graph runtime picks its communicator class (R2, a18).

Inside Ray this module only re-exports Ray's own registry (python/ray/experimental/channel/
accelerator_context.py:19-248), so `register_b200()` is the one line a deployment adds:

    from ant_ray_b200.accelerator_context import register_b200
    register_b200()        # == register_accelerator_context("cuda", B200Communicator)   (:222-233)

after which every default-transport communicator of a compiled graph is built by
`AcceleratorContext.get().create_communicator(world_size, comm_id, rank, actor_handles,
current_stream, use_communication_streams)` (torch_tensor_accelerator_channel.py:673-680) — i.e. as a
B200Communicator.  Without Ray (this image, the tests, bench.py) the same small registry is restated
here with the same names and behaviour, so `channel_context._do_init_communicator` follows the
reference's code path either way.  What cannot be restated without Ray is the GPU assignment lookup
(`ray.get_gpu_ids()`): outside Ray the visible devices are taken from CUDA_VISIBLE_DEVICES / torch.
"""
import importlib
import os
import threading
from contextlib import nullcontext
from typing import List, Optional

try:  # inside Ray: use the real registry so the DAG compiler and this package agree
    from ray.experimental.channel.accelerator_context import (AcceleratorContext, is_accelerator_context_registered,  # noqa: F401
                                                              register_accelerator_context)

    _RAY = True
except ImportError:
    _RAY = False
    _lock = threading.Lock()
    _default_context: Optional["AcceleratorContext"] = None
    _custom_context: Optional["AcceleratorContext"] = None

    class AcceleratorContext:
        """Unified access to streams, events, devices and the communicator class of one accelerator
        backend (reference accelerator_context.py:19-219)."""

        def __init__(self, torch_module_name: str, communicator_cls):
            self._torch_module_name = torch_module_name
            self._communicator_cls = communicator_cls
            if torch_module_name != "cpu":
                self._torch_mod = importlib.import_module(f"torch.{torch_module_name}")

        @staticmethod
        def get() -> "AcceleratorContext":
            """The registered context, else a default chosen from the devices this process sees: the
            B200 communicator on CUDA (the reference picks _NcclGroup there, :64-80)."""
            global _default_context
            with _lock:
                if _custom_context is not None:
                    return _custom_context
                if _default_context is None:
                    import torch

                    if torch.cuda.is_available():
                        from .communicator import B200Communicator

                        _default_context = AcceleratorContext("cuda", B200Communicator)
                    else:
                        _default_context = AcceleratorContext("cpu", None)
                return _default_context

        @staticmethod
        def set(accelerator_context: Optional["AcceleratorContext"]) -> None:
            global _custom_context
            _custom_context = accelerator_context

        def get_accelerator_devices(self) -> List["torch.device"]:
            import torch

            if self._torch_module_name == "cpu":
                return [torch.device("cpu")]
            n = self._torch_mod.device_count()
            # one process per GPU: the process's own (current) device comes first
            cur = self._torch_mod.current_device() if n else 0
            return [torch.device(f"{self._torch_module_name}:{cur}")] if n else [torch.device(f"{self._torch_module_name}:0")]

        def get_device_context(self, device):
            if device.type == "cpu":
                return nullcontext()
            return self._torch_mod.device(device)

        def current_stream(self):
            return self._torch_mod.current_stream()

        def create_event(self):
            return self._torch_mod.Event()

        def generate_communicator_id(self) -> str:
            return self._communicator_cls.generate_communicator_id()

        def create_communicator(self, *args, **kwargs):
            return self._communicator_cls(*args, **kwargs)

        @property
        def module_name(self) -> str:
            return self._torch_module_name

        @property
        def communicator_cls(self):
            return self._communicator_cls

        @property
        def accelerator_count(self) -> int:
            if self._torch_module_name == "cpu":
                return 0
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            if self._torch_module_name == "cuda" and visible is not None:
                return len([v for v in visible.split(",") if v.strip()])
            return self._torch_mod.device_count()

    def register_accelerator_context(torch_module_name: str, communicator_cls) -> None:
        """reference accelerator_context.py:222-233"""
        AcceleratorContext.set(AcceleratorContext(torch_module_name, communicator_cls))

    def is_accelerator_context_registered() -> bool:
        return _custom_context is not None


def register_b200() -> None:
    """Make B200Communicator the communicator of every default-transport ("accelerator") channel and
    collective of compiled graphs in this process."""
    from .communicator import B200Communicator

    register_accelerator_context("cuda", B200Communicator)
