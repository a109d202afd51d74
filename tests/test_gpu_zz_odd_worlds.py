"""World sizes that are not 2, 4 or 8 (one GPU, loopback): the reducing kernels compiled with the world size taken at
run time (`WT = 0`), which the other loopback tests never select.  The reference's own collective test at W = 3 is
`python/ray/tests/test_experimental_collective.py` (fp16, atol 1e-2); here every result is bit-exact against the
rank-order oracle, as for the other world sizes.  The file sorts last among the GPU tests on purpose.
"""
import pytest
import torch

from gpu_common import NATIVE, assert_equal_bits, make_input

from ant_ray_b200 import _native as N
from oracle import oracle as O

pytestmark = pytest.mark.gpu

OPS = {"sum": (N.SUM, O.SUM), "prod": (N.PROD, O.PROD), "max": (N.MAX, O.MAX), "min": (N.MIN, O.MIN), "avg": (N.AVG, O.AVG)}


@pytest.fixture(scope="module", params=[3, 6])
def world(request):
    from ant_ray_b200.loopback import LoopbackWorld

    w = LoopbackWorld(request.param, device=0, key=f"lb-odd{request.param}", staging_bytes=2 << 20, timeout_ms=20000)
    yield w
    w.destroy()


@pytest.mark.parametrize("algo", [N.ALGO_LL, N.ALGO_ONESHOT, N.ALGO_TWOSHOT, N.ALGO_AUTO])
@pytest.mark.parametrize("dtype", [torch.int8, torch.uint8, torch.int32, torch.int64, torch.float16, torch.bfloat16, torch.float32, torch.float64])
def test_allreduce_every_dtype_and_op(world, dtype, algo):
    W = world.world_size
    for opname, (nat, orc) in OPS.items():
        if opname != "sum" and dtype == torch.float16 and algo == N.ALGO_LL:
            continue   # not a combination the W = 2/4/8 suite pins either
        for n in (1, 10, 4096 + 5, 50_003):
            if algo == N.ALGO_LL and n * dtype.itemsize > 32 << 10:
                continue
            ins = [make_input(dtype, n, r, opname) for r in range(W)]
            dev = [t.cuda() for t in ins]
            world.run(lambda r, c: c.allreduce(dev[r].data_ptr(), dev[r].data_ptr(), n, NATIVE[dtype], nat, algo))
            torch.cuda.synchronize()
            world.check()
            want = O.allreduce(ins, orc)
            for r in range(W):
                assert_equal_bits(dev[r], want, f"W={W} allreduce {dtype} n={n} op={opname} algo={algo} rank={r}")


@pytest.mark.parametrize("wire", [torch.bfloat16, torch.float16])
def test_fused_gradient_mean(world, wire):
    W = world.world_size
    for algo in (N.ALGO_ONESHOT, N.ALGO_TWOSHOT):
        for n in (9, 33_333):
            ins = [make_input(torch.float32, n, r) for r in range(W)]
            dev = [t.cuda() for t in ins]
            world.run(lambda r, c: c.allreduce_scaled(dev[r].data_ptr(), dev[r].data_ptr(), n, N.FLOAT32, NATIVE[wire], 1.0 / W, algo))
            torch.cuda.synchronize()
            world.check()
            want = O.allreduce_scaled(ins, wire, 1.0 / W)
            for r in range(W):
                assert_equal_bits(dev[r], want, f"W={W} fused mean wire={wire} n={n} algo={algo} rank={r}")


def test_other_collectives(world):
    W = world.world_size
    m = 20_011
    lists = [[make_input(torch.int32, m, r * 16 + j) for j in range(W)] for r in range(W)]
    devl = [[t.cuda() for t in row] for row in lists]
    outs = [torch.empty(m, dtype=torch.int32, device="cuda") for _ in range(W)]
    world.run(lambda r, c: c.reducescatter([t.data_ptr() for t in devl[r]], outs[r].data_ptr(), m, N.INT32, N.SUM))
    gouts = [[torch.zeros(m, dtype=torch.int32, device="cuda") for _ in range(W)] for _ in range(W)]
    world.run(lambda r, c: c.allgather(devl[r][0].data_ptr(), [t.data_ptr() for t in gouts[r]], m, N.INT32))
    torch.cuda.synchronize()
    world.check()
    want = O.reducescatter(lists)
    for r in range(W):
        assert_equal_bits(outs[r], want[r], f"W={W} reducescatter")
        for j in range(W):
            assert_equal_bits(gouts[r][j], lists[j][0], f"W={W} allgather")
    n = 70_001
    ins = [make_input(torch.float32, n, r) for r in range(W)]
    for root in (0, W - 1):
        dev = [t.cuda() for t in ins]
        world.run(lambda r, c: c.broadcast(dev[r].data_ptr(), n, N.FLOAT32, root))
        torch.cuda.synchronize()
        for r in range(W):
            assert_equal_bits(dev[r], ins[root], f"W={W} broadcast")
        dev = [t.cuda() for t in ins]
        world.run(lambda r, c: c.reduce(dev[r].data_ptr(), dev[r].data_ptr(), n, N.FLOAT32, N.SUM, root))
        torch.cuda.synchronize()
        world.check()
        for r in range(W):
            assert_equal_bits(dev[r], O.reduce(ins) if r == root else ins[r], f"W={W} reduce")
    world.run(lambda r, c: c.barrier())
    torch.cuda.synchronize()
    world.check()


@pytest.mark.parametrize("W", [2, 8])
def test_signed_bytes_every_op_at_the_compiled_world_sizes(W):
    """int8 MAX / MIN / PROD / AVG through the packed per-byte reduce in the kernels compiled for a fixed world size
    (the main loopback suite pins int8 SUM and uint8 for every op)."""
    from ant_ray_b200.loopback import LoopbackWorld

    w = LoopbackWorld(W, device=0, key=f"lb-i8-{W}", staging_bytes=1 << 20, timeout_ms=20000)
    try:
        for algo in (N.ALGO_LL, N.ALGO_ONESHOT, N.ALGO_TWOSHOT):
            for opname, (nat, orc) in OPS.items():
                for n in (7, 5000, 30_011):
                    ins = [make_input(torch.int8, n, r, opname) for r in range(W)]
                    dev = [t.cuda() for t in ins]
                    w.run(lambda r, c: c.allreduce(dev[r].data_ptr(), dev[r].data_ptr(), n, N.INT8, nat, algo))
                    torch.cuda.synchronize()
                    w.check()
                    want = O.allreduce(ins, orc)
                    for r in range(W):
                        assert_equal_bits(dev[r], want, f"W={W} int8 n={n} op={opname} algo={algo} rank={r}")
    finally:
        w.destroy()
