#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 collective / tensor-transport layer.

Metric (BASELINE.json): "allreduce bus GB/s vs msg size; Ray Train ResNet-50 img/s at 1/2/4/8 B200".
  value / e2e        ResNet-50 DDP synthetic-image training throughput (whole job, weak scaling),
                     gradients reduced by the fused peer-memory hook (ant_ray_b200.ddp_hook);
  allreduce_sweep    bus GB/s vs message size, ours next to stock NCCL on the same processes
                     (N >= 2), or two loopback ranks on the one GPU (N = 1);
  roofline           the dominant kernel of OUR path (the fused gradient reduction), timed live with
                     CUDA events on the stream it is launched on;
  cpu_baseline       the reference's CPU path (torch DDP over gloo, which is what Ray Train's
                     _TorchBackend selects without GPUs: train/torch/config.py:167-176) on the
                     box's host cores, bounded sample, N = 1 only.

Launch:  python bench.py --gpus 1 --steps K --warmup W
         python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
             --master-port P bench.py --gpus N --steps K --warmup W
         python bench.py --impl reference ...      (the reference's CPU path, rank 0 only)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

RESNET50_PARAMS = 25_557_032
NVLINK_PEAK_MEASURED = 770.0   # GB/s per direction per GPU, peer copy (B200_PROFILING.md)
NVLINK_PEAK_NOMINAL = 900.0


def log(msg):
    if int(os.environ.get("RANK", 0)) == 0:
        print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--batch", type=int, default=int(os.environ.get("BENCH_BATCH", 256)), help="per-GPU batch")
    p.add_argument("--wire", default=os.environ.get("BENCH_WIRE", "bf16"), choices=["bf16", "fp32", "fp16"])
    p.add_argument("--no-sweep", action="store_true")
    p.add_argument("--no-nccl-ddp", action="store_true")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--sweep-max-bytes", type=int, default=int(os.environ.get("BENCH_SWEEP_MAX", 1 << 30)))
    return p.parse_args()


# ------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi during the timed regions
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.proc = index, [], None
        self.stop_evt = threading.Event()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return self
        threading.Thread(target=self._read, daemon=True).start()
        return self

    def _read(self):
        for line in self.proc.stdout:
            if self.stop_evt.is_set():
                break
            self.samples.append((time.time(), line.strip()))

    def stop(self):
        self.stop_evt.set()
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, windows):
        sm, smax, reasons = [], 0.0, set()
        for ts, line in self.samples:
            if not any(a <= ts <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                smax = max(smax, float(f[1]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# the training workload (reference harness: release/train_tests/benchmark/image_classification/
# factory.py:41 torch.randn(B,3,224,224), :372 torchvision resnet50(weights=None); runner.py:393-403)
# ------------------------------------------------------------------------------------------------
def build_model(device, channels_last=True):
    import torch
    import torchvision

    torch.manual_seed(0)
    model = torchvision.models.resnet50(weights=None)
    model = model.to(device)
    if channels_last and device.type == "cuda":
        model = model.to(memory_format=torch.channels_last)
    return model


def make_step(model, opt, use_autocast, device):
    import torch
    import torch.nn.functional as F

    def step(x, y):
        if use_autocast:
            with torch.autocast(device.type, dtype=torch.bfloat16):
                loss = F.cross_entropy(model(x), y)
        else:
            loss = F.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    return step


def timed_steps(step, x, y, steps, dist, world, pinned=None):
    """Time exactly `steps` steps on the device, barrier + synchronize on both sides, max over ranks.
    With `pinned` = (x_host, y_host) every step copies its inputs from pinned host memory and reads the
    loss back (the end-to-end number)."""
    import torch

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fence()
    t0 = time.time()
    e0.record()
    last = None
    for _ in range(steps):
        if pinned is not None:
            x.copy_(pinned[0], non_blocking=True)
            y.copy_(pinned[1], non_blocking=True)
        loss = step(x, y)
        if pinned is not None:
            last = loss.item()
    e1.record()
    fence()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms, (t0, t1), last


# ------------------------------------------------------------------------------------------------
# allreduce sweep: ours vs NCCL, same processes, same buffers
# ------------------------------------------------------------------------------------------------
def sweep_sizes(max_bytes):
    s, out = 1024, []
    while s <= max_bytes:
        out.append(s)
        s *= 4
    return out


def time_collective(fn, bufs, iters, dist, world, rounds=2):
    """Microseconds per call: every call is bracketed by its own CUDA events (on the launching stream);
    the figure is the median over `iters` calls, best of `rounds` rounds, max over ranks.  Medians and
    a second round keep a transient on the shared host (a ~50 ms slow window was observed once per
    few sweeps, on NCCL and on our kernels alike) from landing in a single size's number."""
    import torch

    best = None
    for _ in range(rounds):
        for i in range(min(5, iters)):
            fn(bufs[i % len(bufs)])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
        for i, (e0, e1) in enumerate(evs):
            e0.record()
            fn(bufs[i % len(bufs)])
            e1.record()
        torch.cuda.synchronize()
        ts = sorted(e0.elapsed_time(e1) * 1e3 for e0, e1 in evs)
        us = ts[len(ts) // 2]
        if world > 1:
            t = torch.tensor([us], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = t.item()
        best = us if best is None else min(best, us)
    return best


def time_fused_bucket(comm, dist, world, wire):
    """The fused gradient kernel alone (full grid, nothing else on the GPU) on ResNet-50's largest bucket
    (30 MiB fp32), in place, rotating over 8 buckets (240 MiB > L2).  Microseconds per launch."""
    import torch

    from ant_ray_b200 import _native as N

    n = 30 << 18
    wire_code = {"bf16": N.BFLOAT16, "fp16": N.FLOAT16, "fp32": N.FLOAT32}[wire]
    bufs = [torch.randn(n, device="cuda") for _ in range(8)]
    us = time_collective(lambda b: comm.allreduce_scaled(b.data_ptr(), b.data_ptr(), n, N.FLOAT32, wire_code, 1.0 / world, N.ALGO_AUTO),
                         bufs, 40, dist, world)
    return us, n


def run_sweep_multi(comm, dist, world, max_bytes):
    """N >= 2: in-place fp32 SUM allreduce of plain torch tensors, ours (AUTO) vs torch c10d NCCL."""
    import torch

    from ant_ray_b200 import _native as N

    rows = []
    for size in sweep_sizes(max_bytes):
        n = size // 4
        nbuf = max(1, min(16, (256 << 20) // size))  # rotate buffers so small sizes are not L2-resident replays
        bufs = [torch.ones(n, dtype=torch.float32, device="cuda") for _ in range(nbuf)]
        iters = 100 if size <= (1 << 20) else (30 if size <= (64 << 20) else 8)
        ours = time_collective(lambda b: comm.allreduce(b.data_ptr(), b.data_ptr(), n, N.FLOAT32, N.SUM, N.ALGO_AUTO), bufs, iters, dist, world)
        nccl = time_collective(lambda b: dist.all_reduce(b), bufs, iters, dist, world)
        k = 2 * (world - 1) / world
        rows.append({"bytes": size, "b200_us": round(ours, 2), "nccl_us": round(nccl, 2),
                     "b200_busbw": round(size / ours / 1e3 * k, 2), "nccl_busbw": round(size / nccl / 1e3 * k, 2)})
        del bufs
    return rows


def run_sweep_loopback(max_bytes):
    """N = 1: two loopback ranks on the one GPU run the same kernels through local HBM."""
    import torch

    from ant_ray_b200 import _native as N
    from ant_ray_b200.loopback import LoopbackWorld

    W = 2
    world = LoopbackWorld(W, device=0, key="bench-sweep", staging_bytes=128 << 20)
    rows = []
    try:
        for size in sweep_sizes(min(max_bytes, 256 << 20)):
            n = size // 4
            bufs = [torch.ones(n, dtype=torch.float32, device="cuda") for _ in range(W)]
            iters = 100 if size <= (1 << 20) else (20 if size <= (64 << 20) else 5)

            def once():
                world.run(lambda r, c: c.allreduce(bufs[r].data_ptr(), bufs[r].data_ptr(), n, N.FLOAT32, N.SUM, N.ALGO_AUTO))

            for _ in range(3):
                once()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                once()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            rows.append({"bytes": size, "b200_us": round(us, 2), "b200_busbw": round(size / us / 1e3, 2), "loopback_ranks": W})
        world.check()
    finally:
        world.destroy()
    return rows


# ------------------------------------------------------------------------------------------------
# reference CPU path: torch DDP over gloo on the host cores (bounded sample)
# ------------------------------------------------------------------------------------------------
def effective_cores():
    """Host cores this container may actually use: affinity mask capped by the cgroup CPU quota."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(quota) // int(period)))
    except (OSError, ValueError):
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                n = min(n, max(1, q // p))
        except (OSError, ValueError):
            pass
    return n


def cpu_has_bf16():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return "avx512_bf16" in flags or "amx_bf16" in flags


def _cpu_worker(rank, world, port, batch, steps, warmup, threads, q, use_bf16, budget_s):
    import torch
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel

    torch.set_num_threads(threads)
    # under torchrun the parent's environment tells c10d to join the launcher's agent store as a client;
    # this private gloo group must bring up its own store instead
    for k in [k for k in os.environ if k.startswith("TORCHELASTIC_") or k in ("GROUP_RANK", "ROLE_RANK", "LOCAL_RANK", "RANK", "WORLD_SIZE",
                                                                                   "LOCAL_WORLD_SIZE", "ROLE_WORLD_SIZE", "GROUP_WORLD_SIZE")]:
        os.environ.pop(k, None)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    device = torch.device("cpu")
    model = DistributedDataParallel(build_model(device, channels_last=False))
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    step = make_step(model, opt, use_autocast=use_bf16, device=device)
    g = torch.Generator().manual_seed(rank)
    x = torch.randn(batch, 3, 224, 224, generator=g)
    y = torch.randint(0, 1000, (batch,), generator=g)
    t_begin = time.time()
    for _ in range(warmup):
        step(x, y)
        if time.time() - t_begin > budget_s / 2:
            break
    dist.barrier()
    t0 = time.time()
    done = 0
    stop = torch.zeros(1)
    for _ in range(steps):
        step(x, y)
        done += 1
        # bounded sample: every rank stops together once the time budget is spent
        stop[0] = 1.0 if time.time() - t_begin > budget_s else 0.0
        dist.all_reduce(stop, op=dist.ReduceOp.MAX)
        if stop.item() > 0:
            break
    dt = time.time() - t0
    if rank == 0:
        q.put((dt, done))
    dist.destroy_process_group()


def cpu_reference(world, batch, steps, warmup, budget_s=60.0):
    """Reference CPU path for this workload: W processes, gloo process group, torch DDP default
    reducer, ResNet-50, bf16 autocast, synthetic images.  Returns (images/s, seconds/step, cores)."""
    import socket

    import torch.multiprocessing as mp

    cores = effective_cores()
    threads = max(1, cores // world)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    use_bf16 = cpu_has_bf16()
    procs = [ctx.Process(target=_cpu_worker, args=(r, world, port, batch, steps, warmup, threads, q, use_bf16, budget_s)) for r in range(world)]
    for p in procs:
        p.start()
    dt, done = q.get(timeout=budget_s * 4 + 120)
    for p in procs:
        p.join(timeout=60)
    return world * batch * done / dt, dt / done, threads * world, done, ("bf16 autocast" if use_bf16 else "fp32 (host CPU has no bf16 units)")


def workload_config(B, world, wire):
    """The workload both arms report (the reference arm runs a bounded sample of it)."""
    return {"workload": "Ray Train TorchTrainer-shaped ResNet-50 DDP step (prepare_model + gradient reduction hook), "
                        "synthetic randn(B,3,224,224), SGD momentum, bf16 autocast, fp32 grads",
            "model": "torchvision.resnet50", "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"dp{world}",
            "grad_wire": wire, "grad_bytes_per_step": RESNET50_PARAMS * 4,
            "l2": "per-step working set (activations of 256 images) is far larger than the 126 MB L2; "
                  "the sweep rotates buffers totalling >= 256 MB"}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    world = max(1, args.gpus)
    batch = int(os.environ.get("BENCH_CPU_BATCH", 16))  # bounded sample: small per-worker batch
    ips, sps, cores, done, cpu_dtype = cpu_reference(world, batch, args.steps, args.warmup, budget_s=float(os.environ.get("BENCH_CPU_BUDGET_S", 150)))
    sample = f"{world} gloo worker(s) x batch {batch}, {done} steps after <= {args.warmup} warm-up, torch DDP default reducer, {cpu_dtype}"
    print(json.dumps({
        "impl": "reference", "metric": "resnet50_ddp_train_images_per_sec", "value": round(ips, 2), "unit": "images/s",
        "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": round(sps * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {**workload_config(args.batch, world, args.wire),
                   "reference_path": "torch DDP default reducer over a gloo process group on the host CPUs (what "
                                     "ray.train.torch.TorchConfig selects without GPUs, train/torch/config.py:167-176)",
                   "sample_per_worker_batch": batch},
        "cpu_baseline": {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(ips, 2), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    from ant_ray_b200 import _native as N
    from ant_ray_b200 import train as b200_train

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun for N > 1"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices (the b200 path has no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    os.environ.setdefault("B200COLL_TIMEOUT_MS", "60000")  # a benchmark should fail fast, not wait out the production default
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=device)
    torch.backends.cudnn.benchmark = True
    N.load()

    B = args.batch
    model = build_model(device)
    model = b200_train.prepare_model(model, grad_wire=args.wire, wrap_single=True)
    state = model.b200_grad_state
    opt = torch.optim.SGD(model.parameters(), lr=0.01, momentum=0.9)
    step = make_step(model, opt, use_autocast=True, device=device)
    g = torch.Generator().manual_seed(1234 + rank)
    x_host = torch.randn(B, 3, 224, 224, generator=g).pin_memory()
    y_host = torch.randint(0, 1000, (B,), generator=g).pin_memory()
    x = x_host.to(device, non_blocking=True).contiguous(memory_format=torch.channels_last)
    y = y_host.to(device, non_blocking=True)
    x_host_cl = x_host.contiguous(memory_format=torch.channels_last).pin_memory()

    log(f"model ready, B={B}, world={world}; warm-up")
    sampler = ClockSampler(local).start() if rank == 0 else None
    for _ in range(max(3, args.warmup)):
        step(x, y)
    log("timing device-resident steps")
    # ---- device-resident inputs
    state.time_kernels = True
    state.events = []
    l0 = N.launch_count()
    if os.environ.get("BENCH_CUDA_PROFILER") == "1":  # ncu --profile-from-start off: capture the timed region only
        torch.cuda.profiler.start()
    ms, win1, _ = timed_steps(step, x, y, args.steps, dist, world)
    if os.environ.get("BENCH_CUDA_PROFILER") == "1":
        torch.cuda.profiler.stop()
    launches = N.launch_count() - l0
    ktimes = state.kernel_times_ms()
    state.time_kernels = False
    log("timing end-to-end steps")
    # ---- end to end: inputs from pinned host memory every step, loss read back every step
    ms_e2e, win2, last_loss = timed_steps(step, x, y, args.steps, dist, world, pinned=(x_host_cl, y_host))
    value = world * B * args.steps / (ms / 1e3)
    e2e = world * B * args.steps / (ms_e2e / 1e3)

    # ---- stock DDP reducer over NCCL on the same box (B-DDP baseline)
    nccl_ddp = None
    log("stock NCCL DDP baseline")
    optional_errors = {}
    if not args.no_nccl_ddp:
        try:  # a failure of an optional section must not cost the headline line
            from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
            from torch.nn.parallel import DistributedDataParallel

            m2 = DistributedDataParallel(build_model(device), device_ids=[device], output_device=device)
            if args.wire == "bf16":
                m2.register_comm_hook(None, default_hooks.bf16_compress_hook)
            o2 = torch.optim.SGD(m2.parameters(), lr=0.01, momentum=0.9)
            s2 = make_step(m2, o2, use_autocast=True, device=device)
            for _ in range(max(3, args.warmup)):
                s2(x, y)
            ms2, _, _ = timed_steps(s2, x, y, args.steps, dist, world)
            nccl_ddp = world * B * args.steps / (ms2 / 1e3)
            del m2, o2, s2
        except Exception as e:  # noqa: BLE001
            optional_errors["nccl_ddp"] = repr(e)[:300]

    # ---- allreduce sweep
    sweep = None
    fused_alone = None
    log("allreduce sweep")
    if not args.no_sweep:
        del model, opt, step
        torch.cuda.empty_cache()
        try:
            if world > 1:
                from ant_ray_b200.b200_group import PeerMemoryComm, next_comm_key

                sweep_comm = PeerMemoryComm(world, rank, next_comm_key("bench-sweep"), local)  # default (full-size) grid
                fused_alone = time_fused_bucket(sweep_comm, dist, world, args.wire)
                sweep = run_sweep_multi(sweep_comm, dist, world, args.sweep_max_bytes)
                sweep_comm.destroy()
            elif rank == 0:
                fused_alone = time_fused_bucket(state.comm, dist, world, args.wire)
                sweep = run_sweep_loopback(args.sweep_max_bytes)
        except Exception as e:  # noqa: BLE001
            optional_errors["allreduce_sweep"] = repr(e)[:300]

    if sampler is not None:
        sampler.stop()
    if rank == 0:
        # ---- roofline of the dominant kernel of our path: the fused reduction of the largest bucket
        wire_b = {"bf16": 2, "fp16": 2, "fp32": 4}[args.wire]
        by_size = {}
        for t_ms, nbytes in ktimes:
            by_size.setdefault(nbytes, []).append(t_ms)
        big = max(by_size) if by_size else 0
        t_big = statistics.mean(by_size[big]) if by_size else None
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        def roofline_for(t_us, nelem, where):
            if world > 1:
                alg = 2 * (world - 1) / world * nelem * wire_b  # NVLink bytes in (== out) per GPU per launch
                ach = alg / (t_us * 1e-6) / 1e9
                return {"bound": "nvlink", "kernel": f"fused gradient allreduce ({args.wire} wire, fp32 accumulate, x1/W), "
                                                     f"{nelem * 4 >> 20} MiB fp32 bucket, {where}",
                        "achieved": round(ach, 1), "peak": NVLINK_PEAK_MEASURED, "peak_nominal": NVLINK_PEAK_NOMINAL, "unit": "GB/s",
                        "frac": round(ach / NVLINK_PEAK_MEASURED, 3), "traffic": None, "launch_us": round(t_us, 2),
                        "algorithmic_bytes": int(alg),
                        "peak_source": "measured peer copy per direction (B200_PROFILING.md), of measured; nominal 900"}
            alg = nelem * 8  # read fp32 + write fp32
            ach = alg / (t_us * 1e-6) / 1e9
            peak = peaks.get("hbm_gbs", 6650.0)
            return {"bound": "hbm", "kernel": f"fused gradient scale / wire rounding (world=1), {nelem * 4 >> 20} MiB fp32 bucket, {where}",
                    "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 3), "traffic": None,
                    "launch_us": round(t_us, 2), "algorithmic_bytes": int(alg),
                    "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)"}

        # `roofline`: the kernel timed alone (what the burst peak is comparable with);
        # `roofline_in_step`: the same kernel inside the training step, where it shares the GPU with the
        # backward pass on a 32-CTA grid and waits for the slowest rank, so it is an upper bound on time.
        roofline_in_step = roofline_for(t_big * 1e3, big // 4, "inside the training step") if t_big else None
        roofline = roofline_for(fused_alone[0], fused_alone[1], "timed alone") if fused_alone else roofline_in_step
        cpu_baseline = None
        if world == 1 and not args.no_cpu_baseline:
            cb = int(os.environ.get("BENCH_CPU_BATCH", 16))
            log("cpu baseline (bounded sample) ...")
            try:
                ips, sps, cores, done, cpu_dtype = cpu_reference(1, cb, 3, 1, budget_s=float(os.environ.get("BENCH_CPU_BUDGET_S", 45)))
                cpu_baseline = {"value": round(ips, 2), "unit": "images/s", "cores": cores, "kind": "port",
                                "sample": f"1 gloo worker x batch {cb}, {done} steps after 1 warm-up (torch DDP default reducer, {cpu_dtype}, host CPU)"}
            except Exception as e:  # noqa: BLE001
                optional_errors["cpu_baseline"] = repr(e)[:300]
        hook_total = sum(t for t, _ in ktimes) / max(1, args.steps)
        out = {
            "metric": "resnet50_ddp_train_images_per_sec", "value": round(value, 1), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(B, world, args.wire),
            "clocks": sampler.summary([win1, win2]) if sampler else None,
            "e2e": {"value": round(e2e, 1), "unit": "images/s", "h2d_bytes_per_step": x_host_cl.numel() * 4 + y_host.numel() * 8,
                    "d2h_bytes_per_step": 4, "ms_per_step": round(ms_e2e / args.steps, 3), "last_loss": last_loss},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "roofline_in_step": roofline_in_step,
            "hook_ms_per_step": round(hook_total, 4),
            "cpu_baseline": cpu_baseline,
            "baselines": {"nccl_ddp_images_per_sec": round(nccl_ddp, 1) if nccl_ddp else None,
                          "nccl_version": ".".join(map(str, torch.cuda.nccl.version()))},
            "multicast": bool(state.comm.multicast),
            "allreduce_sweep": sweep,
        }
        if optional_errors:
            out["optional_section_errors"] = optional_errors
        print(json.dumps(out))
    state.comm.destroy()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
