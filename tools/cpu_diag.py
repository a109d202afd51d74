"""Host-CPU diagnostic for the bench's CPU baseline: cgroup quota, affinity, step time vs threads/dtype."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import build_model, make_step
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    try: print(f, open(f).read().strip())
    except OSError: pass
dev = torch.device("cpu")
for threads in (16, 32, 64):
    for bf16 in (False, True):
        torch.set_num_threads(threads)
        m = build_model(dev, channels_last=False); opt = torch.optim.SGD(m.parameters(), lr=0.01)
        step = make_step(m, opt, bf16, dev)
        x = torch.randn(16, 3, 224, 224); y = torch.randint(0, 1000, (16,))
        t0 = time.time(); step(x, y); t1 = time.time(); step(x, y); t2 = time.time()
        print(f"threads={threads} bf16={bf16}: first {t1-t0:.2f}s second {t2-t1:.2f}s -> {16/(t2-t1):.1f} img/s", flush=True)
        if t2 - t0 > 60: break
