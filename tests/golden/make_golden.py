"""Generate golden vectors by running the reference's CPU arithmetic (real torch.distributed gloo,
driven through the restated TorchGLOOGroup in oracle/gloo_group.py) on seeded inputs.

    python tests/golden/make_golden.py          # rewrites tests/golden/gloo_vectors.pt

The reference itself cannot be imported here (needs a built Ray); what CAN run is the library its
CPU backend delegates every op to (torch_gloo_collective_group.py:147-229).  Torch version used is
recorded in the file.  Inputs are regenerated from seeds by the tests (tests/gpu_common.make_input),
only outputs are stored.
"""
import os
import sys
import tempfile

import torch
import multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORLDS = [2, 3, 4]
N = 37
DTYPES = [torch.int8, torch.uint8, torch.int32, torch.int64, torch.float16, torch.bfloat16, torch.float32, torch.float64]
OPS = ["sum", "prod", "min", "max"]


def worker(rank, world, store_dir, out_q):
    from gpu_common import make_input
    from oracle.gloo_group import GlooGroup
    from ant_ray_b200 import rendezvous, types

    g = GlooGroup(world, rank, "default", 30000, rendezvous.FileStore(store_dir))
    res = {}
    opmap = {"sum": types.ReduceOp.SUM, "prod": types.ReduceOp.PRODUCT, "min": types.ReduceOp.MIN, "max": types.ReduceOp.MAX}
    for dt in DTYPES:
        for op in OPS:
            x = make_input(dt, N, rank, op)
            o = types.AllReduceOptions(); o.reduceOp = opmap[op]
            g.allreduce([x], o)
            res[("allreduce", str(dt), op)] = x.clone()
            x = make_input(dt, N, rank, op)
            o = types.ReduceOptions(); o.reduceOp = opmap[op]; o.root_rank = world - 1
            g.reduce([x], o)
            res[("reduce", str(dt), op)] = x.clone()
        # reducescatter (W allreduces + copy), allgather, broadcast
        lst = [make_input(dt, N, rank * 16 + j) for j in range(world)]
        out = torch.empty_like(lst[0])
        o = types.ReduceScatterOptions(); o.reduceOp = types.ReduceOp.SUM
        g.reducescatter([out], [lst], o)
        res[("reducescatter", str(dt), "sum")] = out.clone()
        x = make_input(dt, N, rank)
        outs = [torch.empty_like(x) for _ in range(world)]
        g.allgather([outs], [x])
        res[("allgather", str(dt), "")] = torch.stack(outs)
        x = make_input(dt, N, rank)
        o = types.BroadcastOptions(); o.root_rank = 1
        g.broadcast([x], o)
        res[("broadcast", str(dt), "")] = x.clone()
    # DDP hook semantics on fp32 gradient buckets (torch default_hooks.py): default = div_(W) then SUM;
    # bf16_compress = to(bf16).div_(W), SUM in bf16, copy back to fp32
    import torch.distributed as dist

    gbuf = make_input(torch.float32, 1000, rank)
    t = gbuf.clone().div_(world); dist.all_reduce(t)
    res[("ddp_default_hook", "torch.float32", "")] = t
    t = gbuf.clone().to(torch.bfloat16).div_(world); dist.all_reduce(t)
    res[("ddp_bf16_compress_hook", "torch.float32", "")] = t.to(torch.float32)
    torch.save(res, os.path.join(store_dir, f"out_{rank}.pt"))
    g.destroy_group()


def main():
    golden = {"torch_version": str(torch.__version__), "n": N, "cases": {}}
    ctx = mp.get_context("spawn")
    for world in WORLDS:
        with tempfile.TemporaryDirectory() as d:
            q = ctx.Queue()
            procs = [ctx.Process(target=worker, args=(r, world, d, q)) for r in range(world)]
            for p in procs:
                p.start()
            for p in procs:
                p.join(timeout=300)
            results = {r: torch.load(os.path.join(d, f"out_{r}.pt")) for r in range(world)}
        golden["cases"][world] = results
    torch.save(golden, os.path.join(HERE, "gloo_vectors.pt"))
    print("wrote", os.path.join(HERE, "gloo_vectors.pt"), os.path.getsize(os.path.join(HERE, "gloo_vectors.pt")), "bytes")


if __name__ == "__main__":
    main()
