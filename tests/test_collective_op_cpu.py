"""Host logic of the compiled-graph collective operation (dag/collective_node.py:176-248) against an
in-process stand-in communicator: output shapes, multi-tensor allreduce (flattened and per-tensor
routes), error cases.  Shapes follow the reference's CPU-communicator DAG test
(dag/tests/experimental/test_cpu_communicator_dag.py:105-134: fp16 fills i + idx, shapes (10*i,))."""
import threading

import pytest
import torch

from ant_ray_b200 import collective_op as C
from ant_ray_b200.types import DagReduceOp


class _Exchange:
    def __init__(self, world):
        self.world, self.lock, self.slots, self.barrier = world, threading.Lock(), {}, threading.Barrier(world)
        self.calls = 0

    def gather(self, key, rank, t):
        with self.lock:
            self.slots.setdefault(key, {})[rank] = t.clone()
        self.barrier.wait(timeout=10)
        parts = [self.slots[key][r] for r in range(self.world)]
        self.barrier.wait(timeout=10)
        return parts


class ThreadCommunicator:
    def __init__(self, rank, ex):
        self.rank, self.ex, self.n = rank, ex, 0

    def _parts(self, t):
        self.n += 1
        return self.ex.gather(self.n, self.rank, t)

    def allreduce(self, s, r, op):
        parts = torch.stack(self._parts(s)).double()
        out = {DagReduceOp.SUM: parts.sum(0), DagReduceOp.PRODUCT: parts.prod(0), DagReduceOp.MAX: parts.max(0).values,
               DagReduceOp.MIN: parts.min(0).values, DagReduceOp.AVG: parts.mean(0)}[op]
        r.copy_(out.to(r.dtype).view(r.shape))

    def allgather(self, s, r):
        r.copy_(torch.cat(self._parts(s)))

    def reducescatter(self, s, r, op):
        total = torch.stack(self._parts(s)).sum(0)
        n = r.shape[0]
        r.copy_(total[self.rank * n:(self.rank + 1) * n])


def run(world, fn):
    ex = _Exchange(world)
    out, errs = [None] * world, []

    def body(r):
        try:
            out[r] = fn(r, ThreadCommunicator(r, ex))
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            ex.barrier.abort()

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]
    return out


def test_allreduce_fp16_fills():
    for i in range(1, 4):
        outs = run(2, lambda r, c: C.CollectiveOperation(c, C.AllReduceOp(), 2).execute(torch.full((10 * i,), float(i + r), dtype=torch.float16)))
        for o in outs:
            assert o.shape == (10 * i,) and (o == 2 * i + 1).all()


def test_allgather_and_reducescatter_shapes():
    outs = run(2, lambda r, c: C.CollectiveOperation(c, C.AllGatherOp(), 2).execute(torch.full((3, 2), float(r))))
    for o in outs:
        assert o.shape == (6, 2) and (o[:3] == 0).all() and (o[3:] == 1).all()
    outs = run(2, lambda r, c: C.CollectiveOperation(c, C.ReduceScatterOp(), 2).execute(torch.arange(8.0).reshape(4, 2) * (r + 1)))
    want = torch.arange(8.0).reshape(4, 2) * 3
    assert torch.equal(outs[0], want[:2]) and torch.equal(outs[1], want[2:])
    with pytest.raises(ValueError):
        C.CollectiveOperation(None, C.ReduceScatterOp(), 2).execute(torch.zeros(3, 2))


@pytest.mark.parametrize("big", [False, True])
def test_multi_tensor_allreduce(big, monkeypatch):
    if big:
        monkeypatch.setattr(C, "FLATTEN_BELOW_BYTES", 0)  # force the per-tensor route
    shapes = [(4,), (2, 3), ()]
    outs = run(2, lambda r, c: C.CollectiveOperation(c, C.AllReduceOp(DagReduceOp.SUM), 2).execute(*[torch.full(s, float(r + k)) for k, s in enumerate(shapes)]))
    for o in outs:
        assert isinstance(o, tuple) and [tuple(t.shape) for t in o] == shapes
        for k, t in enumerate(o):
            assert (t == 2 * k + 1).all()


def test_errors():
    op = C.CollectiveOperation(None, C.AllReduceOp(), 2)
    with pytest.raises(ValueError):
        op.execute(1.0)
    with pytest.raises(ValueError):
        op.execute(torch.zeros(2), torch.zeros(2, dtype=torch.int32))
    with pytest.raises(ValueError):
        C.CollectiveOperation(None, object(), 2).execute(torch.zeros(2))
