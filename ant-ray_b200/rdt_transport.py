"""RDT (GPU objects) tensor transport over the B200 collective group — SURVEY.md section 8(f), N1.

Restates python/ray/experimental/gpu_object_manager/collective_tensor_transport.py:36-184 against the
`TensorTransportManager` contract (tensor_transport_manager.py:14-151): metadata extraction,
communicator lookup through the driver-side registry, and per-tensor send / recv through the
collective API (`collective.send/recv`, reference :128-170).  Differences:
  * `abort_transport` is implemented (the reference raises NotImplementedError): it releases the
    kernels of this actor that wait on the peer.  `can_abort_transport()` nevertheless stays False,
    as in the reference's collective transport: an abort poisons the communicator (sticky abort
    flag, recorded error, the two peers' ring positions no longer agree), so the group cannot carry
    another transfer and Ray must treat the actors as lost, exactly as it does for NCCL;
  * `recv_multiple_tensors` synchronises and checks the communicator before handing the tensors
    to the consumer, so a transfer that failed on the device (peer death, timeout) raises instead of
    returning uninitialised memory.
It is still two-sided (`is_one_sided() == False`): the receiver posts a recv per tensor.
"""
from dataclasses import dataclass, field
from typing import Any, List, Optional, Tuple

from . import collective as _col
from . import experimental_collective as _xc


@dataclass
class TensorTransportMetadata:
    """(shape, dtype) per tensor and the common device (gpu_object_manager/types.py:13-24)."""

    tensor_meta: List[Tuple[Any, Any]] = field(default_factory=list)
    tensor_device: Optional[Any] = None


@dataclass
class CollectiveCommunicatorMetadata:
    communicator_name: str = ""
    src_rank: Optional[int] = None
    dst_rank: Optional[int] = None


class B200TensorTransport:
    def __init__(self, tensor_transport_backend: str = "B200"):
        self._backend = tensor_transport_backend

    @property
    def tensor_transport_backend(self) -> str:
        return self._backend

    @staticmethod
    def is_one_sided() -> bool:
        return False

    @staticmethod
    def can_abort_transport() -> bool:
        return False

    def actor_has_tensor_transport(self, actor) -> bool:
        return len(_xc.get_collective_groups([actor], backend=self._backend)) > 0

    def extract_tensor_transport_metadata(self, obj_id: str, gpu_object: list) -> TensorTransportMetadata:
        meta, device = [], None
        if gpu_object:
            device = gpu_object[0].device
            for t in gpu_object:
                if t.device.type != device.type:
                    raise ValueError("All tensors in an RDT object must have the same device type.")
                meta.append((t.shape, t.dtype))
        return TensorTransportMetadata(tensor_meta=meta, tensor_device=device)

    def get_communicator_metadata(self, src_actor, dst_actor, backend: Optional[str] = None) -> CollectiveCommunicatorMetadata:
        groups = _xc.get_collective_groups([src_actor, dst_actor], backend=backend)
        if len(groups) == 0:
            raise ValueError(f"No communicators found for actors {src_actor} and {dst_actor}. Create a communicator with "
                             "`create_collective_group` before calling actor tasks with non-default tensor_transport.")
        if len(groups) > 1:
            raise ValueError(f"There are {len(groups)} possible communicators that contain actors {src_actor} and {dst_actor}. "
                             "Currently, RDT objects only support one communicator.")
        g = groups[0]
        src, dst = g.get_rank(src_actor), g.get_rank(dst_actor)
        if src == -1 or dst == -1:
            raise ValueError("Sender and receiver must be in the same communicator.")
        return CollectiveCommunicatorMetadata(communicator_name=g.name, src_rank=src, dst_rank=dst)

    def recv_multiple_tensors(self, tensors: list, obj_id: str, tensor_transport_metadata: TensorTransportMetadata,
                              communicator_metadata: CollectiveCommunicatorMetadata):
        assert isinstance(communicator_metadata, CollectiveCommunicatorMetadata)
        for t in tensors:
            _col.recv(t, communicator_metadata.src_rank, communicator_metadata.communicator_name)
        g = _col.get_group_handle(communicator_metadata.communicator_name)
        check = getattr(g, "check", None)
        if tensors and check is not None:
            check(synchronize=True)  # the data must have landed, and landed intact, before the consumer sees it

    def send_multiple_tensors(self, tensors: list, tensor_transport_metadata: TensorTransportMetadata,
                              communicator_metadata: CollectiveCommunicatorMetadata):
        device = tensors[0].device if tensors else None
        for t in tensors:
            if t.device.type != device.type:
                raise ValueError(f"tensor device {t.device} does not match device {device}")
            _col.send(t, communicator_metadata.dst_rank, communicator_metadata.communicator_name)

    def garbage_collect(self, obj_id: str, tensor_transport_meta: TensorTransportMetadata):
        pass  # nothing is registered per object: staging lives in the communicator's arena

    def abort_transport(self, obj_id: str, communicator_metadata: CollectiveCommunicatorMetadata):
        """Release any kernel of this actor that is waiting on the peer (inside the actor process)."""
        g = _col.get_group_handle(communicator_metadata.communicator_name)
        comm = getattr(g, "_comm", None)
        if comm is not None:
            comm.abort()
