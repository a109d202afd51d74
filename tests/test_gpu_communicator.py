"""R2 on real GPUs (>= 2): B200Communicator + TorchTensor channel between two actor processes.

Follows the reference's GPU DAG tests (python/ray/dag/tests/experimental/test_torch_tensor_dag.py):
p2p with varying shapes (:213-268), dedicated communication streams (:373-415), collectives for
every reduce op compared with torch.equal against torch.sum/prod/min/max of the stacked inputs
(:1340-1450, W=2), wrong-shape allreduce surfaces RayChannelError instead of hanging (:1544-1588),
destroyed group raises RayChannelError; and the microbenchmark shape of
release/microbenchmark/experimental/compiled_graph_gpu_microbenchmark.py (fp16, 100,000 bytes).
"""
import multiprocessing as mp

import pytest
import torch

from mini_actor import get, spawn

pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


class CommActor:
    def __init__(self, rank, world, store_dir, overlap=False):
        import os

        os.environ["B200COLL_STORE"] = f"file://{store_dir}"
        os.environ.setdefault("B200COLL_TIMEOUT_MS", "10000")
        torch.cuda.set_device(rank)
        self.rank, self.world, self.overlap = rank, world, overlap
        self.comm = None
        self.chan = None

    def connect(self, comm_id):
        from ant_ray_b200.communicator import B200Communicator

        self.comm = B200Communicator(self.world, comm_id, self.rank, list(range(self.world)), torch.cuda.current_stream(), self.overlap)
        return self.comm.get_self_rank()

    def open_channel(self, conn, writer, readers, static_shape=False, direct=False, cpu_conn=None):
        from ant_ray_b200.channel import PipeMetaChannel, TensorListChannel, TorchTensorChannel

        inner = TensorListChannel(self.comm, writer, readers, PipeMetaChannel(conn), static_shape)
        self.chan = TorchTensorChannel(inner, PipeMetaChannel(cpu_conn), direct) if cpu_conn is not None else inner
        self.meta = inner._meta
        return True

    def send_tensors(self, specs, fill):
        ts = [torch.full(shape, fill + i, dtype=dtype, device="cuda") for i, (shape, dtype) in enumerate(specs)]
        with self.comm.send_stream:
            self.chan.write(ts)
        torch.cuda.synchronize()
        return self.meta.writes

    def recv_tensors(self):
        with self.comm.recv_stream:
            ts = self.chan.read(timeout=30)
        torch.cuda.synchronize()
        return [t.cpu() for t in ts], self.meta.reads

    def send_value(self, value_kind):
        v = {"t": torch.arange(12, device="cuda").reshape(3, 4), "tag": "hello", "n": 3} if value_kind == "dict" else torch.ones(5, device="cuda") * 7
        self.chan.write(v)
        torch.cuda.synchronize()
        return True

    def recv_value(self):
        v = self.chan.read(timeout=30)
        torch.cuda.synchronize()
        return {"t": v["t"].cpu(), "tag": v["tag"], "n": v["n"]} if isinstance(v, dict) else v.cpu()

    def collective(self, kind, op_name, n, dtype, seed_shift=0, numel_override=None):
        from ant_ray_b200.types import DagReduceOp

        g = torch.Generator().manual_seed(100 + self.rank + seed_shift)
        x = torch.randn(n if numel_override is None else numel_override, generator=g).to(dtype).cuda()
        op = getattr(DagReduceOp, op_name)
        if kind == "allreduce":
            out = torch.empty_like(x)
            self.comm.allreduce(x, out, op)
        elif kind == "allgather":
            out = torch.empty(x.numel() * self.world, dtype=dtype, device="cuda")
            self.comm.allgather(x, out)
        else:
            out = torch.empty(x.numel() // self.world, dtype=dtype, device="cuda")
            self.comm.reducescatter(x, out, op)
        return out.cpu()

    def try_after_destroy(self):
        from ant_ray_b200.communicator import RayChannelError

        self.comm.destroy()
        try:
            self.comm.send(torch.ones(4, device="cuda"), 1 - self.rank)
        except RayChannelError:
            return "RayChannelError"
        return "no error"

    def mismatch(self, n):
        from ant_ray_b200.communicator import RayChannelError

        x = torch.ones(n, device="cuda")
        try:
            self.comm.allreduce(x, torch.empty_like(x))
            self.comm.check()   # errors are deferred (no host sync in the collective): check() surfaces them
        except RayChannelError as e:
            return "RayChannelError: " + str(e)[:80]
        return "no error"

    def p2p_latency(self, nbytes, iters, sender):
        """Microbenchmark shape: fp16 vector of `nbytes`; returns microseconds per message."""
        import time

        n = nbytes // 2
        t = torch.ones(n, dtype=torch.float16, device="cuda")
        alloc = lambda shape, dtype: torch.empty(shape, dtype=dtype, device="cuda")  # noqa: E731
        for phase in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(iters):
                if sender:
                    self.comm.send(t, 1 - self.rank)
                else:
                    self.comm.recv((n,), torch.float16, 1 - self.rank, alloc)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        return dt / iters * 1e6

    def close(self):
        if self.comm is not None:
            self.comm.destroy()
        return True


@pytest.fixture
def pair(store_dir):
    if _ngpu() < 2:
        pytest.skip("needs >= 2 GPUs")
    made = []

    def make(overlap=False):
        actors = [spawn(CommActor, r, 2, store_dir, overlap, start_method="spawn") for r in range(2)]
        made.extend(actors)
        assert get([a.connect.remote("cg-" + str(len(made))) for a in actors]) == [0, 1]
        return actors

    yield make
    for a in made:
        try:
            get(a.close.remote(), timeout=20)
        except Exception:
            pass
        a.kill()


@pytest.mark.parametrize("overlap", [False, True])
def test_p2p_varying_shapes_and_static_shape(pair, overlap):
    a0, a1 = pair(overlap)
    ctx = mp.get_context("spawn")
    c0, c1 = ctx.Pipe()
    get([a0.open_channel.remote(c0, 0, [1]), a1.open_channel.remote(c1, 0, [1])])
    for i, shape in enumerate([(10,), (20, 3), (1,), (4, 5, 6)]):
        specs = [(shape, torch.float16), ((3,), torch.int64)]
        refs = [a0.send_tensors.remote(specs, float(i)), a1.recv_tensors.remote()]
        writes, (got, reads) = get(refs)
        assert writes == 0 and reads == 0  # dynamic shapes travel in the communicator's header ring, not the side channel
        assert torch.equal(got[0], torch.full(shape, float(i), dtype=torch.float16))
        assert torch.equal(got[1], torch.full((3,), i + 1, dtype=torch.int64))


def test_static_shape_and_direct_return_skip_cpu_hops(pair):
    a0, a1 = pair()
    ctx = mp.get_context("spawn")
    c0, c1 = ctx.Pipe()
    get([a0.open_channel.remote(c0, 0, [1], True), a1.open_channel.remote(c1, 0, [1], True)])
    for i in range(4):
        writes, (got, reads) = get([a0.send_tensors.remote([((50_000,), torch.float16)], float(i)), a1.recv_tensors.remote()])
        assert writes == 0 and reads == 0  # headers are inlined; after the first message not even those are sent
        assert (got[0] == i).all()


def test_value_channel_with_cpu_remainder(pair):
    a0, a1 = pair()
    ctx = mp.get_context("spawn")
    m0, m1 = ctx.Pipe()
    p0, p1 = ctx.Pipe()
    get([a0.open_channel.remote(m0, 0, [1], False, False, p0), a1.open_channel.remote(m1, 0, [1], False, False, p1)])
    _, got = get([a0.send_value.remote("dict"), a1.recv_value.remote()])
    assert torch.equal(got["t"], torch.arange(12).reshape(3, 4)) and got["tag"] == "hello" and got["n"] == 3


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_collectives_match_torch(pair, dtype):
    a0, a1 = pair()
    n = 2 * 3000
    ins = [torch.randn(n, generator=torch.Generator().manual_seed(100 + r)).to(dtype) for r in range(2)]
    stacked = torch.stack(ins)
    expect = {"SUM": stacked.sum(0), "PRODUCT": stacked.prod(0), "MIN": stacked.min(0).values, "MAX": stacked.max(0).values,
              "AVG": (stacked.float().sum(0) / 2).to(dtype)}
    for op, want in expect.items():
        outs = get([a.collective.remote("allreduce", op, n, dtype) for a in (a0, a1)])
        for o in outs:
            assert torch.equal(o, want), op  # W=2 is order-independent: exact (test_torch_tensor_dag.py:1340-1450)
    outs = get([a.collective.remote("allgather", "SUM", n, dtype) for a in (a0, a1)])
    for o in outs:
        assert torch.equal(o, torch.cat(ins))
    outs = get([a.collective.remote("reducescatter", "SUM", n, dtype) for a in (a0, a1)])
    for r, o in enumerate(outs):
        assert torch.equal(o, stacked.sum(0)[r * (n // 2):(r + 1) * (n // 2)])


def test_wrong_shape_raises_channel_error_not_hang(pair):
    a0, a1 = pair()
    # beyond the LL size both ranks compare signatures in the flag round; in the LL range the rank with the longer
    # message is guaranteed to notice (the shorter one receives all it waits for and checks best-effort)
    res = get([a0.mismatch.remote(100_000), a1.mismatch.remote(200_000)], timeout=60)
    assert all(r.startswith("RayChannelError") for r in res), res


def test_destroy_raises_channel_error(pair):
    a0, a1 = pair()
    assert get(a0.try_after_destroy.remote()) == "RayChannelError"


def test_microbenchmark_shape_100kB_fp16(pair):
    """compiled_graph_gpu_microbenchmark.py:441-451: 100,000-byte fp16 tensor, sender -> receiver."""
    a0, a1 = pair()
    us = get([a0.p2p_latency.remote(100_000, 200, True), a1.p2p_latency.remote(100_000, 200, False)])
    print(f"p2p 100kB fp16: sender {us[0]:.1f} us/msg, receiver {us[1]:.1f} us/msg")
    assert us[1] < 2000
