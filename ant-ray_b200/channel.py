"""TorchTensor accelerator channel over a Communicator (R2).

Restates the data path of python/ray/experimental/channel/torch_tensor_accelerator_channel.py:
the inner `_TorchTensorAcceleratorChannel` (:368-649) — metadata (shape, dtype) over a CPU side
channel, then one communicator.send per tensor per reader (:549-590) / one communicator.recv per
tensor (:610-641), with `_static_shape` skipping the metadata hop after the first message
(:485-547, :592-608) — and the outer channel's split of a value into out-of-band GPU tensors plus a
pickled remainder (:183-305, serialization_context.py:94-168), with `_direct_return` skipping the
CPU hop entirely (:260-274).

Ray's shared-memory `Channel` is the metadata side channel in production; anything with
`write(obj)` / `read(timeout)` works (tests and the 2-actor benchmark use a multiprocessing pipe).
"""
import pickle
from dataclasses import dataclass
from typing import Any, List, Optional, Tuple


@dataclass(frozen=True)
class TorchTensorMetadata:
    """Shape and dtype a reader needs to allocate the receive buffer (reference :41-47)."""

    shape: Tuple[int, ...]
    dtype: Any


def default_allocator(shape, dtype):
    import torch

    return torch.empty(tuple(shape), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))


class PipeMetaChannel:
    """Metadata side channel over multiprocessing connections: one writer end, N reader ends."""

    def __init__(self, conns):
        self.conns = conns if isinstance(conns, (list, tuple)) else [conns]
        self.writes = 0
        self.reads = 0

    def write(self, obj, timeout=None):
        self.writes += 1
        for c in self.conns:
            c.send(obj)

    def read(self, timeout=None):
        c = self.conns[0]
        if timeout is not None and not c.poll(timeout):
            raise TimeoutError("timed out waiting for tensor metadata")
        self.reads += 1
        return c.recv()

    def close(self):
        for c in self.conns:
            c.close()


class TensorListChannel:
    """Lists of GPU tensors: metadata, then payload over the communicator.

    Two things the reference leaves as TODOs are used when the communicator offers them
    (B200Communicator does; the CPU test double and foreign communicators do not):
      * `inline_metadata`: shape and dtype ride in a binary header next to the communicator's cell ring
        (`send_with_header` / `recv_with_header`), so dynamic shapes need no second channel and no
        pickle (reference :574-578, :592-608); `meta_channel` is then only used if inlining is refused
        (`inline_metadata=False`);
      * `multi_reader`: with several readers every tensor is sent ONCE (`send_multi`, a multicast store
        stream) instead of once per reader (reference :586-590, "TODO: ... can replace with a
        broadcast").  A writer can multicast to one reader set only; further channels of the same
        writer with other readers must be created with `multicast=False`.
    Both choices are made identically on the writer and on every reader from the constructor
    arguments, which the creator of the channel passes to all endpoints alike."""

    def __init__(self, communicator, writer_rank: int, reader_ranks: List[int], meta_channel, static_shape: bool = False,
                 allocator=default_allocator, require_cuda: bool = True, inline_metadata: bool = True, multicast: bool = True):
        self._comm = communicator
        self._writer_rank, self._reader_ranks = writer_rank, list(reader_ranks)
        self._meta = meta_channel
        self._static_shape = static_shape
        self._static_meta: Optional[List[TorchTensorMetadata]] = None
        self._allocator = allocator
        self._require_cuda = require_cuda  # False only with the CPU test double of the communicator
        me = communicator.get_self_rank()
        self._is_writer = me == writer_rank
        self._is_reader = me in self._reader_ranks
        self._inline = bool(inline_metadata and getattr(communicator, "inline_metadata", False))
        self._multi = bool(multicast and len(self._reader_ranks) > 1 and getattr(communicator, "multi_reader", False))

    def _send_metadata(self, tensors):
        import torch

        meta = []
        for t in tensors:
            if not isinstance(t, torch.Tensor):
                raise ValueError("Task must return torch.Tensors")
            if self._require_cuda and not t.is_cuda:
                raise ValueError("torch.Tensor must be on the default (cuda) device")
            meta.append(TorchTensorMetadata(tuple(t.shape), t.dtype))
        if self._static_meta is not None:
            if meta != self._static_meta:
                raise ValueError(f"Expected torch.Tensors with shapes and dtypes: {self._static_meta}, found: {meta}. DAG will shut down.")
            return None  # the readers already know: no metadata hop
        if self._static_shape:
            self._static_meta = meta
        return meta

    def write(self, tensors: List["torch.Tensor"], timeout: Optional[float] = None):
        assert self._is_writer, "this actor is not the writer of the channel"
        meta = self._send_metadata(tensors)
        with_header = meta is not None and self._inline
        if meta is not None and not self._inline:
            self._meta.write(meta)  # before the sends, so the reader can launch the matching recv
        if with_header and not tensors:
            self._comm.announce_empty(self._reader_ranks)
        for i, t in enumerate(tensors):
            if self._multi:
                if with_header:
                    self._comm.send_with_header(t, self._reader_ranks, i, len(tensors))
                else:
                    self._comm.send_multi(t, self._reader_ranks)
                continue
            for rank in self._reader_ranks:
                if with_header:
                    self._comm.send_with_header(t, rank, i, len(tensors))
                else:
                    self._comm.send(t, rank)

    def read(self, timeout: Optional[float] = None) -> List["torch.Tensor"]:
        assert self._is_reader, "this actor is not a reader of the channel"
        meta = self._static_meta
        if meta is None and self._inline:
            out, metas = [], []
            t, _, count = self._comm.recv_with_header(self._writer_rank, self._allocator, timeout, multi=self._multi)
            for i in range(count):
                if i:
                    t, index, cnt = self._comm.recv_with_header(self._writer_rank, self._allocator, timeout, multi=self._multi)
                    if index != i or cnt != count:
                        raise ValueError(f"tensor header out of sequence: got {index}/{cnt}, expected {i}/{count}")
                out.append(t)
                metas.append(TorchTensorMetadata(tuple(t.shape), t.dtype))
            if self._static_shape:
                self._static_meta = metas
            return out
        if meta is None:
            meta = self._meta.read(timeout)
            if self._static_shape:
                self._static_meta = meta
        recv = self._comm.recv_multi if self._multi else self._comm.recv
        return [recv(m.shape, m.dtype, self._writer_rank, self._allocator) for m in meta]

    def close(self):
        self._meta.close()
        self._comm.destroy()


class _Placeholder:
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index


class _TensorExtractingPickler(pickle.Pickler):
    """Replace CUDA tensors by integer placeholders while pickling the rest of the value
    (reference serialization_context.py:94-111)."""

    def __init__(self, file, tensors, require_cuda=True):
        super().__init__(file, protocol=pickle.HIGHEST_PROTOCOL)
        self.tensors = tensors
        self.require_cuda = require_cuda

    def persistent_id(self, obj):
        import torch

        if isinstance(obj, torch.Tensor) and (obj.is_cuda or not self.require_cuda):
            self.tensors.append(obj)
            return len(self.tensors) - 1
        return None


class _TensorRestoringUnpickler(pickle.Unpickler):
    def __init__(self, file, tensors):
        super().__init__(file)
        self.tensors = tensors

    def persistent_load(self, pid):
        return self.tensors[pid]


class TorchTensorChannel:
    """Arbitrary Python values whose CUDA tensors travel out of band over the communicator
    (reference outer channel :49-352).  `direct_return`: the value IS one CUDA tensor, no CPU hop."""

    def __init__(self, tensor_channel: TensorListChannel, cpu_channel, direct_return: bool = False):
        self._gpu = tensor_channel
        self._cpu = cpu_channel
        self._direct_return = direct_return

    def write(self, value, timeout: Optional[float] = None):
        import io

        import torch

        if self._direct_return:
            if not (isinstance(value, torch.Tensor) and (value.is_cuda or not self._gpu._require_cuda)):
                raise ValueError("Task annotated with _direct_return=True must return a CUDA torch.Tensor, "
                                 f"instead found value `{type(value).__name__}`. DAG will shut down.")
            self._gpu.write([value], timeout)
            return
        tensors: List[torch.Tensor] = []
        f = io.BytesIO()
        _TensorExtractingPickler(f, tensors, self._gpu._require_cuda).dump(value)
        # tensors first: the reader posts its receives, then reads the CPU remainder
        self._gpu.write(tensors, timeout)
        self._cpu.write(f.getvalue())

    def read(self, timeout: Optional[float] = None):
        import io

        tensors = self._gpu.read(timeout)
        if self._direct_return:
            return tensors[0]
        data = self._cpu.read(timeout)
        return _TensorRestoringUnpickler(io.BytesIO(data), tensors).load()

    def close(self):
        self._gpu.close()
        self._cpu.close()
