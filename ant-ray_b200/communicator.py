"""B200Communicator: the compiled-graph accelerator communicator (R2), replacing _NcclGroup.

`Communicator` restates the 15-method contract of
python/ray/experimental/channel/communicator.py:18-199; inside a Ray installation the real ABC is
used instead so that `isinstance` checks in the DAG compiler hold.  `B200Communicator` keeps the
constructor the registry calls — `Cls(world_size, comm_id, rank, actor_handles, cuda_stream,
use_communication_streams)` (torch_tensor_accelerator_channel.py:673-680) — and the behaviour of
python/ray/experimental/channel/nccl_group.py:21-374:
  * send() returns once the kernel is enqueued (:149-184);
  * recv() allocates through the caller's allocator, then blocks the host until the data has
    landed and raises RayChannelError if the group was destroyed meanwhile (:186-241);
  * allgather/allreduce/reducescatter are out of place, synchronise, and raise RayChannelError when
    the group was closed or a peer disagreed on the shape/dtype (:243-333; the reference relies on
    an NCCL timeout there, test_torch_tensor_dag.py:1544-1588 — here the kernels compare op
    signatures and fail in microseconds);
  * destroy() sets `_closed` first, then aborts the in-flight kernels (:347-365).
Register it with `register_accelerator_context("cuda", B200Communicator)`
(accelerator_context.py:222-233) or pass an instance as `transport=`.
"""
import uuid
from abc import ABC, abstractmethod
from typing import Callable, Optional, Tuple

from . import _native as N
from .b200_group import PeerMemoryComm, TensorView, native_reduce_op
from .types import DagReduceOp as ReduceOp

try:  # inside Ray: be a real subclass so the DAG compiler's isinstance checks pass
    from ray.exceptions import RayChannelError
    from ray.experimental.channel.communicator import Communicator
except ImportError:

    class RayChannelError(RuntimeError):
        """Raised when a channel / communicator is closed or failed (ray.exceptions.RayChannelError)."""

    class Communicator(ABC):
        """Contract of ray.experimental.channel.Communicator (communicator.py:18-199)."""

        @abstractmethod
        def initialize(self, rank: int) -> None: ...

        @abstractmethod
        def get_actor_handles(self) -> list: ...

        @abstractmethod
        def get_rank(self, actor) -> int: ...

        @abstractmethod
        def get_self_rank(self) -> Optional[int]: ...

        def get_world_size(self) -> int:
            raise NotImplementedError

        @abstractmethod
        def send(self, value, peer_rank: int) -> None: ...

        @abstractmethod
        def recv(self, shape, dtype, peer_rank: int, allocator=None): ...

        @property
        @abstractmethod
        def recv_stream(self): ...

        @property
        @abstractmethod
        def send_stream(self): ...

        @abstractmethod
        def allgather(self, send_buf, recv_buf) -> None: ...

        @abstractmethod
        def allreduce(self, send_buf, recv_buf, op) -> None: ...

        @abstractmethod
        def reducescatter(self, send_buf, recv_buf, op) -> None: ...

        @abstractmethod
        def destroy(self) -> None: ...

        @abstractmethod
        def get_transport_name(self) -> str: ...

        @classmethod
        @abstractmethod
        def generate_communicator_id(cls) -> str: ...


TorchTensorAllocator = Callable[[Tuple[int], "torch.dtype"], "torch.Tensor"]


def _actor_key(a):
    return getattr(a, "_ray_actor_id", a)


class B200Communicator(Communicator):
    """One actor's endpoint of a peer-memory group.  Not thread-safe (like _NcclGroup)."""

    def __init__(self, world_size: int, comm_id: str, rank: Optional[int], actor_handles: list,
                 cuda_stream: Optional["torch.cuda.Stream"], use_communication_streams: bool = False,
                 store=None, config=None):
        self._world_size = world_size
        self._rank = rank
        self._actor_handles = actor_handles
        self._use_communication_streams = use_communication_streams
        self._comm: Optional[PeerMemoryComm] = None
        self._cuda_stream = self._send_stream = self._recv_stream = None
        self._closed = False
        if rank is not None:
            import torch

            assert cuda_stream is not None, "B200 actor must specify cuda_stream"
            assert torch.cuda.is_available(), "B200 actor has no GPUs assigned"
            device = cuda_stream.device.index if hasattr(cuda_stream, "device") else torch.cuda.current_device()
            # blocks until the same call has been made on every other actor of the group
            self._comm = PeerMemoryComm(world_size, rank, f"b200coll/cgraph/{comm_id}", device, store, config)
            self._cuda_stream = cuda_stream
            if use_communication_streams:
                self._send_stream = torch.cuda.Stream(device=device)
                self._recv_stream = torch.cuda.Stream(device=device)
            else:
                self._send_stream = self._recv_stream = cuda_stream

    # -- membership ---------------------------------------------------------------------------
    def initialize(self, rank: int) -> None:
        pass  # everything happens in the constructor, as in _NcclGroup

    def get_actor_handles(self) -> list:
        return self._actor_handles

    def get_rank(self, actor) -> int:
        keys = [_actor_key(a) for a in self._actor_handles]
        try:
            return keys.index(_actor_key(actor))
        except ValueError:
            raise ValueError("Actor is not in the B200 group.")

    def get_self_rank(self) -> Optional[int]:
        return self._rank

    def get_world_size(self) -> int:
        return self._world_size

    def get_transport_name(self) -> str:
        return "accelerator"

    @classmethod
    def generate_communicator_id(cls) -> str:
        return uuid.uuid4().hex

    # -- helpers ------------------------------------------------------------------------------
    def _check_open(self):
        if self._closed or self._comm is None or self._comm.handle is None:
            raise RayChannelError("B200 group has been destroyed.")

    def _raise_if_failed(self, what: str):
        if self._closed:
            raise RayChannelError(f"B200 group has been destroyed during {what}.")
        try:
            self._comm.check()
        except N.B200CollError as e:
            raise RayChannelError(f"B200 {what} failed: {e}. There may be a shape or dtype mismatch between "
                                  "the tensors of different ranks, or a peer actor died.") from e

    # -- p2p ----------------------------------------------------------------------------------
    def send(self, buf, peer_rank: int) -> None:
        self._check_open()
        if self._use_communication_streams:
            # keep the CPU loop from running arbitrarily far ahead of the GPU (nccl_group.py:168-173)
            self._send_stream.synchronize()
        v = TensorView(buf)
        try:
            self._comm.send(v.ptr, v.numel * v.itemsize, peer_rank, stream=self._send_stream)
        except N.B200CollError as e:
            raise RayChannelError(str(e)) from e

    def recv(self, shape, dtype, peer_rank: int, allocator: Optional[TorchTensorAllocator] = None):
        self._check_open()
        assert allocator is not None, "B200 group requires a tensor allocator"
        buf = allocator(shape, dtype)
        if self._use_communication_streams:
            self._recv_stream.synchronize()
        v = TensorView(buf)
        try:
            self._comm.recv(v.ptr, v.numel * v.itemsize, peer_rank, stream=self._recv_stream)
        except N.B200CollError as e:
            raise RayChannelError(str(e)) from e
        if not self._use_communication_streams:
            # "After this call returns, the receive buffer is safe to read from any stream"
            self._recv_stream.synchronize()
            self._raise_if_failed("recv")
        return buf

    # -- collectives --------------------------------------------------------------------------
    def _exec_collective(self, send_buf, recv_buf, what, fn):
        self._check_open()
        assert send_buf.dtype == recv_buf.dtype, (
            "Ray Compiled Graph derived the dtype of recv_buf from send_buf, so send_buf and recv_buf must have the same dtype.")
        import torch

        try:
            with torch.cuda.stream(self._cuda_stream):
                fn()
        except N.B200CollError as e:
            raise RayChannelError(str(e)) from e
        self._cuda_stream.synchronize()
        self._raise_if_failed(what)

    def allgather(self, send_buf, recv_buf) -> None:
        s, r = TensorView(send_buf), TensorView(recv_buf)
        if r.numel != s.numel * self._world_size:
            raise RayChannelError(f"allgather recv_buf has {r.numel} elements, expected {s.numel * self._world_size}")
        step = s.numel * s.itemsize
        self._exec_collective(send_buf, recv_buf, "allgather",
                              lambda: self._comm.allgather(s.ptr, [r.ptr + j * step for j in range(self._world_size)], s.numel, s.dtype))

    def allreduce(self, send_buf, recv_buf, op=ReduceOp.SUM) -> None:
        s, r = TensorView(send_buf), TensorView(recv_buf)
        if r.numel != s.numel:
            raise RayChannelError(f"allreduce recv_buf has {r.numel} elements, expected {s.numel}")
        self._exec_collective(send_buf, recv_buf, "allreduce",
                              lambda: self._comm.allreduce(s.ptr, r.ptr, s.numel, s.dtype, native_reduce_op(op)))

    def reducescatter(self, send_buf, recv_buf, op=ReduceOp.SUM) -> None:
        s, r = TensorView(send_buf), TensorView(recv_buf)
        if s.numel != r.numel * self._world_size:
            raise RayChannelError(f"reducescatter send_buf has {s.numel} elements, expected {r.numel * self._world_size}")
        step = r.numel * r.itemsize
        self._exec_collective(send_buf, recv_buf, "reducescatter",
                              lambda: self._comm.reducescatter([s.ptr + j * step for j in range(self._world_size)], r.ptr, r.numel,
                                                               r.dtype, native_reduce_op(op)))

    # -- streams / teardown -------------------------------------------------------------------
    @property
    def recv_stream(self):
        import torch

        return torch.cuda.StreamContext(self._recv_stream)

    @property
    def send_stream(self):
        import torch

        return torch.cuda.StreamContext(self._send_stream)

    def destroy(self) -> None:
        if self._closed:
            return
        self._closed = True  # before the abort, so ops released by it see the flag (nccl_group.py:355-363)
        if self._comm is not None:
            self._comm.abort()
            self._comm.destroy()
