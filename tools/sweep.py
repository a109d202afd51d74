"""Collective micro-benchmark: ours (per algorithm) next to stock NCCL, same processes, same buffers.

Launch under torchrun:  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/sweep.py [opts]
  --algos auto,oneshot,twoshot,nvls,nvls_sym   which of our paths to time
  --sizes 1024,...    bytes (default 1 KB .. 1 GB x4)
  --dtype f32|bf16
  --ops allreduce,allgather,reducescatter,broadcast,sendrecv
Prints one JSON object (rank 0).  nccl-tests conventions: algBW = S/t, busBW = algBW * 2(W-1)/W (allreduce).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ant_ray_b200 import _native as N  # noqa: E402
from ant_ray_b200.b200_group import PeerMemoryComm, make_config  # noqa: E402


PER_ITER = False
LAST_STATS = {}


def timeit(fn, bufs, iters, world):
    """Mean microseconds per call over a back-to-back loop (nccl-tests style), max over ranks.
    With --per-iter every call is bracketed by its own events and LAST_STATS gets min/median/max."""
    for i in range(min(5, iters)):
        fn(bufs[i % len(bufs)])
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    if PER_ITER:
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (a, b) in enumerate(evs):
            a.record(); fn(bufs[i % len(bufs)]); b.record()
        torch.cuda.synchronize()
        ts = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
        t = torch.tensor([ts[len(ts) // 2], ts[0], ts[-1], sum(ts) / len(ts)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        LAST_STATS.update(median=round(t[0].item(), 2), min=round(t[1].item(), 2), max=round(t[2].item(), 2))
        return t[3].item()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(bufs[i % len(bufs)])
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) * 1e3 / iters], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algos", default="auto,oneshot,twoshot,nvls,nvls_sym")
    ap.add_argument("--sizes", default="")
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--ops", default="allreduce")
    ap.add_argument("--max-blocks", type=int, default=0)
    ap.add_argument("--staging-mb", type=int, default=256)
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--per-iter", action="store_true", help="time every call separately; report min/median/max")
    ap.add_argument("--blocks-list", default="", help="extra communicators with these max_blocks, e.g. 592,1184")
    ap.add_argument("--variants", default="", help="extra communicators with config overrides, timed on one algorithm each: "
                    "'label:algo:key=val,key=val;label2:algo2:...' e.g. 'sym64:nvls_sym:nvls_blocks=64;r16:nvls_pipe:granule_bytes=16384'")
    ap.add_argument("--iters-scale", type=float, default=1.0)
    ap.add_argument("--wire", default="", choices=["", "bf16", "f16"], help="f32 buffers with a 16-bit wire type: the fused gradient-mean call (allreduce_scaled)")
    a = ap.parse_args()
    global PER_ITER
    PER_ITER = a.per_iter
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dtype = {"f32": torch.float32, "bf16": torch.bfloat16}[a.dtype]
    nat = {"f32": N.FLOAT32, "bf16": N.BFLOAT16}[a.dtype]
    esz = 4 if a.dtype == "f32" else 2
    sizes = [int(s) for s in a.sizes.split(",")] if a.sizes else [1024 * 4**k for k in range(11) if 1024 * 4**k <= a.max_bytes]
    kw = dict(staging_bytes=a.staging_mb << 20, symmetric_bytes=max(sizes) if "nvls_sym" in a.algos else 0)
    if a.max_blocks:
        kw["max_blocks"] = a.max_blocks
    comm = PeerMemoryComm(world, rank, "sweep", local, None, make_config(**kw))
    extra = {}
    for mb in [int(x) for x in a.blocks_list.split(",") if x]:
        extra[mb] = PeerMemoryComm(world, rank, f"sweep-mb{mb}", local, None, make_config(**{**kw, "max_blocks": mb}))
    variants = []
    for spec in [v for v in a.variants.split(";") if v]:
        label, algo_name, *rest = spec.split(":")
        over = dict(kw)
        for item in (rest[0].split(",") if rest and rest[0] else []):
            k_, v_ = item.split("=")
            over[k_] = int(v_)
        if algo_name == "nvls_sym":
            over["symmetric_bytes"] = max(sizes)
        variants.append((label, algo_name, PeerMemoryComm(world, rank, f"sweep-var-{label}", local, None, make_config(**over))))
    algos = {"auto": N.ALGO_AUTO, "ll": N.ALGO_LL, "nvls_lanes": N.ALGO_NVLS_LANES, "nvls_streams": N.ALGO_NVLS_STREAMS, "oneshot": N.ALGO_ONESHOT, "twoshot": N.ALGO_TWOSHOT, "nvls": N.ALGO_NVLS, "nvls_sym": N.ALGO_NVLS,
             "nvls_pipe": N.ALGO_NVLS_PIPE}
    out = {"world": world, "dtype": a.dtype, "multicast": bool(comm.multicast), "nccl_version": ".".join(map(str, torch.cuda.nccl.version())),
           "max_blocks": comm.config.max_blocks, "rows": []}
    k = 2 * (world - 1) / world
    wire_nat = {"": None, "bf16": N.BFLOAT16, "f16": N.FLOAT16}[a.wire]

    def ar(cx, b, algo):
        if wire_nat is None:
            cx.allreduce(b.data_ptr(), b.data_ptr(), n, nat, N.SUM, algo)
        else:
            cx.allreduce_scaled(b.data_ptr(), b.data_ptr(), n, nat, wire_nat, 1.0 / world, algo)

    for op in a.ops.split(","):
        for size in sizes:
            n = size // esz
            nbuf = max(1, min(16, (256 << 20) // size))
            iters = max(3, int((200 if size <= (1 << 20) else (40 if size <= (64 << 20) else 10)) * a.iters_scale))
            row = {"op": op, "bytes": size}
            if op == "allreduce":
                bufs = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(nbuf)]
                for label, algo_name, cx in variants:
                    if algo_name.startswith("nvls") and not cx.multicast:
                        continue
                    if algo_name == "ll" and size > cx.config.ll_max_bytes:
                        continue
                    if algo_name == "nvls_sym":
                        vb = [cx.symmetric_tensor((n,), dtype)]
                        us = timeit(lambda b: cx.allreduce(b.data_ptr(), b.data_ptr(), n, nat, N.SUM, N.ALGO_NVLS), vb, iters, world)
                    else:
                        us = timeit(lambda b: ar(cx, b, algos[algo_name]), bufs, iters, world)
                    row[label + "_us"] = round(us, 2)
                    row[label + "_busbw"] = round(size / us / 1e3 * k, 1)
                for name in [x for x in a.algos.split(",") if x]:
                    if name in ("nvls", "nvls_sym", "nvls_pipe") and not comm.multicast:
                        continue
                    if name == "ll" and size > comm.config.ll_max_bytes:
                        continue
                    if name == "oneshot" and size * world > (a.staging_mb << 20) * 4:
                        continue
                    if name == "nvls_sym":
                        sb = [comm.symmetric_tensor((n,), dtype)]
                        us = timeit(lambda b: comm.allreduce(b.data_ptr(), b.data_ptr(), n, nat, N.SUM, N.ALGO_NVLS), sb, iters, world)
                    else:
                        us = timeit(lambda b: ar(comm, b, algos[name]), bufs, iters, world)
                    row[name + "_us"] = round(us, 2)
                    row[name + "_busbw"] = round(size / us / 1e3 * k, 1)
                    if PER_ITER:
                        row[name + "_stats"] = dict(LAST_STATS)
                for mb, cx in extra.items():
                    for name in ("twoshot", "nvls"):
                        if name not in a.algos.split(",") or (name == "nvls" and not cx.multicast):
                            continue
                        us = timeit(lambda b: cx.allreduce(b.data_ptr(), b.data_ptr(), n, nat, N.SUM, algos[name]), bufs, iters, world)
                        row[f"{name}_mb{mb}_us"] = round(us, 2); row[f"{name}_mb{mb}_busbw"] = round(size / us / 1e3 * k, 1)
                if not a.no_nccl:
                    us = timeit(lambda b: dist.all_reduce(b), bufs, iters, world)
                    row["nccl_us"] = round(us, 2); row["nccl_busbw"] = round(size / us / 1e3 * k, 1)
            elif op == "allgather":
                bufs = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(nbuf)]
                outs = [torch.empty(n, dtype=dtype, device="cuda") for _ in range(world)]
                flat = torch.empty(n * world, dtype=dtype, device="cuda")
                ptrs = [o.data_ptr() for o in outs]
                us = timeit(lambda b: comm.allgather(b.data_ptr(), ptrs, n, nat), bufs, iters, world)
                row["b200_us"] = round(us, 2); row["b200_busbw"] = round(size * world / us / 1e3 * (world - 1) / world, 1)
                if not a.no_nccl:
                    # what the reference does: allGather into a flat buffer, then W copies (nccl_collective_group.py:278-296)
                    def ref(b):
                        dist.all_gather_into_tensor(flat, b)
                        for j in range(world):
                            outs[j].copy_(flat[j * n:(j + 1) * n])
                    us = timeit(ref, bufs, iters, world)
                    row["nccl_ref_us"] = round(us, 2); row["nccl_ref_busbw"] = round(size * world / us / 1e3 * (world - 1) / world, 1)
            elif op == "reducescatter":
                lists = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(world)]
                o = torch.empty(n, dtype=dtype, device="cuda")
                flat = torch.empty(n * world, dtype=dtype, device="cuda")
                ptrs = [t.data_ptr() for t in lists]
                us = timeit(lambda b: comm.reducescatter(ptrs, o.data_ptr(), n, nat, N.SUM), [None], iters, world)
                row["b200_us"] = round(us, 2); row["b200_busbw"] = round(size * world / us / 1e3 * (world - 1) / world, 1)
                if not a.no_nccl:
                    def ref(b):
                        for j in range(world):
                            flat[j * n:(j + 1) * n].copy_(lists[j])
                        dist.reduce_scatter_tensor(o, flat)
                    us = timeit(ref, [None], iters, world)
                    row["nccl_ref_us"] = round(us, 2); row["nccl_ref_busbw"] = round(size * world / us / 1e3 * (world - 1) / world, 1)
            elif op == "broadcast":
                bufs = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(nbuf)]
                us = timeit(lambda b: comm.broadcast(b.data_ptr(), n, nat, 0), bufs, iters, world)
                row["b200_us"] = round(us, 2); row["b200_busbw"] = round(size / us / 1e3, 1)
                if not a.no_nccl:
                    us = timeit(lambda b: dist.broadcast(b, 0), bufs, iters, world)
                    row["nccl_us"] = round(us, 2); row["nccl_busbw"] = round(size / us / 1e3, 1)
            elif op == "sendrecv":
                bufs = [torch.ones(n, dtype=dtype, device="cuda") for _ in range(nbuf)]
                def ours(b):
                    if rank == 0: comm.send(b.data_ptr(), size, 1)
                    elif rank == 1: comm.recv(b.data_ptr(), size, 0)
                us = timeit(ours, bufs, iters, world)
                row["b200_us"] = round(us, 2); row["b200_gbps"] = round(size / us / 1e3, 1)
                if not a.no_nccl:
                    def ref(b):
                        if rank == 0: dist.send(b, 1)
                        elif rank == 1: dist.recv(b, 0)
                    us = timeit(ref, bufs, iters, world)
                    row["nccl_us"] = round(us, 2); row["nccl_gbps"] = round(size / us / 1e3, 1)
            out["rows"].append(row)
            if rank == 0:
                print("#", json.dumps(row), flush=True)
    comm.check()
    for cx in list(extra.values()) + [v[2] for v in variants]:
        cx.check(); cx.destroy()
    if rank == 0:
        print(json.dumps(out))
    comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
