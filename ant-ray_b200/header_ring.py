"""Tensor headers (shape, dtype) next to the p2p cell ring — SURVEY.md section 8(f) N3.

The reference tells the reader what to allocate through a second channel: a pickled list of
`_TorchTensorMetadata` over Ray's shared-memory mutable-object channel, one semaphore hand-shake per
message (torch_tensor_accelerator_channel.py:574-578 writer, :592-608 reader;
shared_memory_channel.py:435-499).  After the GPU hop got fast that hop dominates a 100 kB message.

Here the header travels in a fixed-size binary record — no pickle, no semaphore, no system call in
steady state — through a single-writer / single-reader ring in a shared-memory file that belongs to
the communicator, one ring per ordered (writer, reader) pair, written by the sender's host right
before it enqueues the send kernel and polled by the receiver's host right before it allocates and
enqueues the receive.  The record carries the per-pair message number, so a header can never be
matched with the wrong payload: record k announces exactly the k-th message of that pair's cell ring.

Layout of the file (little endian):
    [0]   u64 written      records published so far (writer only)
    [8]   u64 consumed     records consumed so far  (reader only)
    [64]  SLOTS records of 128 bytes:
            u64 seq (= record number + 1, written LAST), u32 crc of the rest, u16 index, u16 count,
            u16 dtype code, u16 ndim, i64 shape[8], padding
Both sides validate seq and crc, so a torn read (the ring is lock-free) is retried, not trusted.
"""
import mmap
import os
import struct
import time
import zlib
from typing import Optional, Tuple

SLOTS = 64
RECORD = 128
_HDR = 64
_FILE_BYTES = _HDR + SLOTS * RECORD
_BODY = struct.Struct("<HHHH8q")      # index, count, dtype, ndim, shape[8]
_MAX_DIMS = 8

_DTYPE_CODES = None


def _codes():
    global _DTYPE_CODES
    if _DTYPE_CODES is None:
        import torch

        names = ["uint8", "int8", "int16", "int32", "int64", "float16", "bfloat16", "float32", "float64", "bool",
                 "uint16", "uint32", "uint64", "complex64", "complex128", "float8_e4m3fn", "float8_e5m2"]
        pairs = [(getattr(torch, n), i + 1) for i, n in enumerate(names) if hasattr(torch, n)]
        _DTYPE_CODES = (dict(pairs), {c: d for d, c in pairs})
    return _DTYPE_CODES


class HeaderTimeout(TimeoutError):
    pass


class HeaderRing:
    """One direction of one pair.  `role` is "w" (the sender's process) or "r" (the receiver's)."""

    def __init__(self, path: str, role: str):
        assert role in ("w", "r")
        self.path, self.role = path, role
        fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o600)
        try:
            if os.fstat(fd).st_size < _FILE_BYTES:
                os.ftruncate(fd, _FILE_BYTES)   # new pages read as zero: written = consumed = 0, every seq = 0
            self.mm = mmap.mmap(fd, _FILE_BYTES)
        finally:
            os.close(fd)
        self.count = 0   # records written (writer) / consumed (reader) by THIS endpoint

    # -- writer -------------------------------------------------------------------------------------
    def put(self, shape: Tuple[int, ...], dtype, index: int = 0, count: int = 1, timeout_s: float = 60.0):
        if len(shape) > _MAX_DIMS:
            raise ValueError(f"tensor headers carry at most {_MAX_DIMS} dimensions, got {len(shape)}")
        code = _codes()[0].get(dtype)
        if code is None:
            raise ValueError(f"dtype {dtype} has no header code")
        k = self.count
        deadline = None
        while k - struct.unpack_from("<Q", self.mm, 8)[0] >= SLOTS:   # ring full: the reader is SLOTS messages behind
            if deadline is None:
                deadline = time.monotonic() + timeout_s
            elif time.monotonic() > deadline:
                raise HeaderTimeout("tensor-header ring is full: the reader stopped consuming")
            time.sleep(0)
        body = _BODY.pack(index, count, code, len(shape), *(list(shape) + [0] * (_MAX_DIMS - len(shape))))
        off = _HDR + (k % SLOTS) * RECORD
        self.mm[off + 12:off + 12 + len(body)] = body
        struct.pack_into("<I", self.mm, off + 8, zlib.crc32(body + struct.pack("<Q", k + 1)))
        struct.pack_into("<Q", self.mm, off, k + 1)          # publish: seq last
        self.count = k + 1
        struct.pack_into("<Q", self.mm, 0, self.count)

    # -- reader -------------------------------------------------------------------------------------
    def get(self, timeout_s: Optional[float] = 60.0, cancelled=None):
        """Blocks (spinning, then yielding) until record number `self.count` is there; returns
        (shape, dtype, index, count).  `cancelled()` is polled so that a destroy() from another thread
        releases a reader."""
        k = self.count
        off = _HDR + (k % SLOTS) * RECORD
        deadline = None if timeout_s is None else time.monotonic() + timeout_s
        spins = 0
        while True:
            if struct.unpack_from("<Q", self.mm, off)[0] == k + 1:
                body = bytes(self.mm[off + 12:off + 12 + _BODY.size])
                if struct.unpack_from("<I", self.mm, off + 8)[0] == zlib.crc32(body + struct.pack("<Q", k + 1)):
                    index, count, code, ndim, *shape = _BODY.unpack(body)
                    self.count = k + 1
                    struct.pack_into("<Q", self.mm, 8, self.count)
                    return tuple(shape[:ndim]), _codes()[1][code], index, count
            spins += 1
            if spins > 2000:
                if cancelled is not None and cancelled():
                    raise HeaderTimeout("cancelled")
                if deadline is not None and time.monotonic() > deadline:
                    raise HeaderTimeout(f"no tensor header (message {k}) within {timeout_s}s")
                time.sleep(0 if spins < 20000 else 0.0002)

    def close(self, unlink: bool = False):
        if unlink and self.role == "w":
            # the reader opens its end lazily: a writer that unlinks unread records would strand it
            try:
                unlink = struct.unpack_from("<Q", self.mm, 8)[0] >= self.count
            except ValueError:
                unlink = False
        try:
            self.mm.close()
        except (BufferError, ValueError):
            pass
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def ring_path(comm_key: str, src: int, dst: int) -> str:
    """Every endpoint derives the same file name from the communicator's rendezvous key."""
    import hashlib

    root = os.environ.get("B200COLL_HEADER_DIR", "/dev/shm")
    return os.path.join(root, "b200coll-hdr-" + hashlib.sha1(comm_key.encode()).hexdigest()[:20] + f"-{src}-{dst}")
