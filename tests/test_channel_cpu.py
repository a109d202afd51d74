"""Host logic of the TorchTensor channel (R2) with a CPU test double of the communicator.

Mirrors the reference's mocked-NCCL channel tests (python/ray/tests/test_nccl_channel.py:222-300:
`TracedChannel` counts metadata writes to prove that `_static_shape` / `_direct_return` skip the
metadata and CPU hops) and the serialization round trips of
python/ray/tests/test_channel_serialization.py.  The arithmetic-free double moves tensors through
queues; the real communicator is covered by the GPU tests.
"""
import queue

import pytest
import torch

from ant_ray_b200.channel import PipeMetaChannel, TensorListChannel, TorchTensorChannel, TorchTensorMetadata
from ant_ray_b200.communicator import B200Communicator, Communicator, RayChannelError


class QueueCommunicator(Communicator):
    """In-process stand-in: send() enqueues, recv() dequeues and checks shape/dtype."""

    def __init__(self, rank, queues, actors=("a0", "a1", "a2")):
        self.rank, self.queues, self.actors = rank, queues, list(actors)
        self.sent = 0

    def initialize(self, rank): pass
    def get_actor_handles(self): return self.actors
    def get_rank(self, actor): return self.actors.index(actor)
    def get_self_rank(self): return self.rank
    def get_world_size(self): return len(self.actors)

    def send(self, value, peer_rank):
        self.sent += 1
        self.queues[(self.rank, peer_rank)].put(value.clone())

    def recv(self, shape, dtype, peer_rank, allocator=None):
        t = self.queues[(peer_rank, self.rank)].get(timeout=5)
        if tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            raise RayChannelError("shape/dtype mismatch")
        buf = allocator(shape, dtype)
        buf.copy_(t)
        return buf

    recv_stream = send_stream = property(lambda self: None)
    def allgather(self, s, r): raise NotImplementedError
    def allreduce(self, s, r, op): raise NotImplementedError
    def reducescatter(self, s, r, op): raise NotImplementedError
    def destroy(self): pass
    def get_transport_name(self): return "accelerator"
    @classmethod
    def generate_communicator_id(cls): return "q"


class ListMeta:
    def __init__(self):
        self.q, self.writes, self.reads = queue.Queue(), 0, 0

    def write(self, obj, timeout=None):
        self.writes += 1
        self.q.put(obj)

    def read(self, timeout=None):
        self.reads += 1
        return self.q.get(timeout=5)

    def close(self): pass


def cpu_alloc(shape, dtype):
    return torch.empty(tuple(shape), dtype=dtype)


def make_pair(static_shape=False, readers=(1,)):
    queues = {(0, r): queue.Queue() for r in readers}
    meta = {r: ListMeta() for r in readers}

    class Fan:  # writer side fans the metadata out to every reader's side channel
        writes = 0

        def write(self, obj, timeout=None):
            Fan.writes += 1
            for m in meta.values():
                m.write(obj)

        def close(self): pass

    w = TensorListChannel(QueueCommunicator(0, queues), 0, list(readers), Fan(), static_shape, cpu_alloc, require_cuda=False)
    rs = [TensorListChannel(QueueCommunicator(r, queues), 0, list(readers), meta[r], static_shape, cpu_alloc, require_cuda=False) for r in readers]
    return w, rs, Fan, meta


def test_dynamic_shapes_send_metadata_every_time():
    w, (r,), fan, meta = make_pair()
    for shape in [(10,), (3, 4), (0,), (2, 2, 2)]:
        t = torch.randn(shape)
        w.write([t])
        (got,) = r.read()
        assert torch.equal(got, t)
    assert fan.writes == 4 and meta[1].reads == 4


def test_static_shape_skips_metadata_after_first_message():
    w, (r,), fan, meta = make_pair(static_shape=True)
    for i in range(5):
        t = torch.full((7,), float(i), dtype=torch.float16)
        w.write([t])
        assert torch.equal(r.read()[0], t)
    assert fan.writes == 1 and meta[1].reads == 1
    with pytest.raises(ValueError):  # a different shape on a static-shape channel is an error at the writer
        w.write([torch.zeros(8, dtype=torch.float16)])
    with pytest.raises(ValueError):
        w.write([torch.zeros(7, dtype=torch.float32)])


def test_multiple_tensors_and_readers():
    w, rs, fan, _ = make_pair(readers=(1, 2))
    a, b = torch.arange(6).reshape(2, 3), torch.ones(4, dtype=torch.bfloat16)
    w.write([a, b])
    for r in rs:
        ga, gb = r.read()
        assert torch.equal(ga, a) and torch.equal(gb, b)
    assert w._comm.sent == 4  # one send per tensor per reader (reference :586-590)


def test_non_tensor_is_rejected():
    w, _, _, _ = make_pair()
    with pytest.raises(ValueError):
        w.write([1.0])


@pytest.mark.parametrize("direct", [False, True])
def test_outer_channel_roundtrip(direct):
    w, (r,), fan, meta = make_pair()
    cpu_w, cpu_r = ListMeta(), None
    cw = TorchTensorChannel(w, cpu_w, direct_return=direct)
    cr = TorchTensorChannel(r, cpu_w, direct_return=direct)
    if direct:
        t = torch.randn(5)
        cw.write(t)
        assert torch.equal(cr.read(), t)
        assert cpu_w.writes == 0  # no CPU hop at all
        with pytest.raises(ValueError):
            cw.write({"not": "a tensor"})
    else:
        value = {"x": torch.randn(3), "meta": ("step", 7), "nested": [torch.zeros(()), "s"]}
        cw.write(value)
        got = cr.read()
        assert torch.equal(got["x"], value["x"]) and got["meta"] == ("step", 7)
        assert torch.equal(got["nested"][0], torch.zeros(())) and got["nested"][1] == "s"
        cw.write("no tensors at all")
        assert cr.read() == "no tensors at all"


def test_scalar_and_dtypes_roundtrip():
    w, (r,), _, _ = make_pair()
    for dt in (torch.float16, torch.bfloat16, torch.float32, torch.int64, torch.uint8, torch.bool):
        for shape in [(), (3,), (2, 0, 4)]:
            t = torch.zeros(shape, dtype=dt)
            w.write([t])
            got = r.read()[0]
            assert got.dtype == dt and got.shape == t.shape


def test_metadata_dataclass_equality():
    assert TorchTensorMetadata((2, 3), torch.float16) == TorchTensorMetadata((2, 3), torch.float16)
    assert TorchTensorMetadata((2, 3), torch.float16) != TorchTensorMetadata((3, 2), torch.float16)


def test_b200_communicator_driver_side_and_membership():
    """rank=None is the driver's handle (nccl_group.py:96-98): no GPU work, membership queries only."""
    actors = ["a", "b"]
    c = B200Communicator(2, "id", None, actors, None)
    assert c.get_self_rank() is None and c.get_world_size() == 2
    assert c.get_rank("b") == 1 and c.get_actor_handles() == actors
    with pytest.raises(ValueError):
        c.get_rank("zzz")
    assert c.get_transport_name() == "accelerator"
    assert isinstance(B200Communicator.generate_communicator_id(), str)
    with pytest.raises(RayChannelError):
        c.send(torch.ones(2), 1)
    c.destroy()
    c.destroy()  # idempotent
    with pytest.raises(RayChannelError):
        c.recv((2,), torch.float32, 0, cpu_alloc)


def test_pipe_meta_channel():
    import multiprocessing as mp

    a, b = mp.Pipe()
    w, r = PipeMetaChannel(a), PipeMetaChannel(b)
    w.write([TorchTensorMetadata((1, 2), torch.int32)])
    assert r.read(1) == [TorchTensorMetadata((1, 2), torch.int32)]
    with pytest.raises(TimeoutError):
        r.read(0.01)


class InlineCommunicator(QueueCommunicator):
    """CPU double of a communicator that, like B200Communicator, carries tensor headers itself (N3) and can
    deliver one payload to several readers (N2).  Headers and payloads travel through the same queues."""

    inline_metadata = True
    multi_reader = True

    def __init__(self, rank, queues, mqueues):
        super().__init__(rank, queues)
        self.mq = mqueues       # (src, dst) -> queue of multi-reader payloads
        self.multi_sends = 0
        self.headers = 0

    def send_multi(self, value, peer_ranks):
        self.multi_sends += 1
        for p in peer_ranks:
            self.mq[(self.rank, p)].put(value.clone())

    def recv_multi(self, shape, dtype, peer_rank, allocator=None):
        t = self.mq[(peer_rank, self.rank)].get(timeout=5)
        assert tuple(t.shape) == tuple(shape) and t.dtype == dtype
        buf = allocator(shape, dtype)
        buf.copy_(t)
        return buf

    def send_with_header(self, buf, peer_ranks, index=0, count=1):
        peers = [peer_ranks] if isinstance(peer_ranks, int) else list(peer_ranks)
        for p in peers:
            self.headers += 1
            self.queues[(self.rank, p)].put(("hdr", tuple(buf.shape), buf.dtype, index, count))
        if len(peers) == 1:
            self.send(buf, peers[0])
        else:
            self.send_multi(buf, peers)

    def announce_empty(self, peer_ranks):
        for p in ([peer_ranks] if isinstance(peer_ranks, int) else list(peer_ranks)):
            self.headers += 1
            self.queues[(self.rank, p)].put(("hdr", (), torch.uint8, 0, 0))

    def recv_with_header(self, peer_rank, allocator=None, timeout=None, multi=False):
        tag, shape, dtype, index, count = self.queues[(peer_rank, self.rank)].get(timeout=5)
        assert tag == "hdr"
        if count == 0:
            return None, 0, 0
        fn = self.recv_multi if multi else self.recv
        return fn(shape, dtype, peer_rank, allocator), index, count


def make_inline(readers=(1,), static_shape=False, **kw):
    queues = {(0, r): queue.Queue() for r in readers}
    mqueues = {(0, r): queue.Queue() for r in readers}
    side = ListMeta()
    w = TensorListChannel(InlineCommunicator(0, queues, mqueues), 0, list(readers), side, static_shape, cpu_alloc, require_cuda=False, **kw)
    rs = [TensorListChannel(InlineCommunicator(r, queues, mqueues), 0, list(readers), side, static_shape, cpu_alloc, require_cuda=False, **kw)
          for r in readers]
    return w, rs, side


def test_inline_headers_replace_the_metadata_channel():
    """N3: with a communicator that carries headers, dynamic shapes never touch the side channel
    (reference torch_tensor_accelerator_channel.py:574-578, 592-608)."""
    w, (r,), side = make_inline()
    for shape in [(10,), (3, 4), (0,), (2, 2, 2)]:
        a, b = torch.randn(shape), torch.arange(3)
        w.write([a, b])
        ga, gb = r.read()
        assert torch.equal(ga, a) and torch.equal(gb, b)
    w.write([])                      # an empty list still announces itself
    assert r.read() == []
    assert side.writes == 0 and side.reads == 0 and w._comm.headers == 9


def test_inline_static_shape_sends_headers_once():
    w, (r,), side = make_inline(static_shape=True)
    for i in range(4):
        t = torch.full((7,), float(i))
        w.write([t])
        assert torch.equal(r.read()[0], t)
    assert w._comm.headers == 1 and side.writes == 0
    with pytest.raises(ValueError):
        w.write([torch.zeros(8)])


def test_multi_reader_channel_sends_each_tensor_once():
    """N2: one send_multi per tensor instead of one send per reader (reference :586-590 TODO)."""
    w, rs, side = make_inline(readers=(1, 2, 3))
    a, b = torch.arange(6).reshape(2, 3), torch.ones(4, dtype=torch.bfloat16)
    w.write([a, b])
    for r in rs:
        ga, gb = r.read()
        assert torch.equal(ga, a) and torch.equal(gb, b)
    assert w._comm.multi_sends == 2 and w._comm.sent == 0 and w._comm.headers == 6
    # static shapes: later messages are bare multi-sends
    w2, rs2, _ = make_inline(readers=(1, 2), static_shape=True)
    for i in range(3):
        w2.write([torch.full((5,), float(i))])
        for r in rs2:
            assert torch.equal(r.read()[0], torch.full((5,), float(i)))
    assert w2._comm.multi_sends == 3 and w2._comm.headers == 2


def test_multicast_and_inlining_can_be_switched_off_per_channel():
    w, rs, side = make_inline(readers=(1, 2), multicast=False, inline_metadata=False)
    t = torch.randn(6)
    w.write([t])
    # the side channel carries the metadata once per reader read, the payload goes out once per reader
    for r in rs:
        side.q.put([TorchTensorMetadata((6,), torch.float32)]) if side.q.empty() else None
        assert torch.equal(r.read()[0], t)
    assert w._comm.sent == 2 and w._comm.multi_sends == 0 and w._comm.headers == 0 and side.writes == 1
