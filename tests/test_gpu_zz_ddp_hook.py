"""The fused gradient hook INSIDE torch DDP on one GPU (R3).

`prepare_model(..., wrap_single=True)` wraps the model in DistributedDataParallel at world size 1 and registers
`b200_allreduce_hook`; every bucket then goes through the library (`k_local_scale*`: wire rounding + scale, the
W = 1 form of the fused reduction).  The hook is wrapped so that each bucket is also reduced the stock way on a
copy — torch's default reducer arithmetic for the fp32 wire, `bf16_compress_hook`'s for the bf16 wire — and every
parameter's `.grad` must hold exactly those bits afterwards.  Then a few optimizer steps run through the hooked
model to exercise the futures / stream hand-over repeatedly.

The multi-GPU form of this check (W = 2, 4, 8 against NCCL) is the `parity.ddp_grads_*` block of `bench.py`.
The file sorts last among the GPU tests on purpose.  It runs in a subprocess: DDP needs a default process group,
which must not leak into the other tests of the session.
"""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
sys.path.insert(0, os.environ["B200_TEST_ROOT"])
import torch
import torch.distributed as dist
import torch.nn as nn

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
from ant_ray_b200 import ddp_hook
from ant_ray_b200 import train as b200_train


def build():
    torch.manual_seed(7)
    # 0.5 MiB + 4 MiB + 40 KiB of fp32 parameters; with bucket_cap_mb=1 DDP usually cuts them into one bucket above and
    # one below the hook's small-bucket threshold (communication-stream path / producer-stream path).  The expectations
    # below hold for whatever bucketing torch chooses.
    return nn.Sequential(nn.Linear(128, 1024), nn.ReLU(), nn.Linear(1024, 1024), nn.ReLU(), nn.Linear(1024, 10))


def flat(t):
    return t.as_strided((t.numel(),), (1,), t.storage_offset())


g = torch.Generator().manual_seed(11)
x = torch.randn(64, 128, generator=g).to(dev)
y = torch.randint(0, 10, (64,), generator=g).to(dev)
fused_hook = ddp_hook.b200_allreduce_hook
for wire in ("fp32", "bf16"):
    expected, sizes = {}, []

    def both(st, bucket):
        buf = bucket.buffer()
        nbytes = buf.numel() * buf.element_size()
        sizes.append(nbytes)
        if wire == "fp32" or nbytes <= ddp_hook.SMALL_BUCKET_BYTES:
            ref = buf.clone().div_(1)                       # default reducer: buffer / W, allreduce
        else:
            ref = buf.to(torch.bfloat16).div_(1).float()    # bf16_compress_hook: cast, / W, allreduce, copy back
        for p_, gview in zip(bucket.parameters(), bucket.gradients()):
            off = gview.storage_offset() - buf.storage_offset()
            expected[p_] = ref[off:off + gview.numel()]
        return fused_hook(st, bucket)

    ddp_hook.b200_allreduce_hook = both     # what prepare_model -> ddp_hook.register attaches (DDP takes one hook only)
    model = b200_train.prepare_model(build(), grad_wire=wire, wrap_single=True, parallel_strategy_kwargs={"bucket_cap_mb": 1})
    ddp_hook.b200_allreduce_hook = fused_hook
    state = model.b200_grad_state
    loss = nn.functional.cross_entropy(model(x), y)
    loss.backward()
    torch.cuda.synchronize()
    state.comm.check()
    params = [p_ for p_ in model.parameters() if p_.grad is not None]
    assert len(params) == 6 and len(expected) == 6, (len(params), len(expected))
    assert len(sizes) >= 1 and sum(sizes) == 4 * sum(p_.numel() for p_ in params), sizes
    assert state.launches == len(sizes), (state.launches, sizes)
    for p_ in params:
        a, b = flat(p_.grad), expected[p_]
        assert torch.equal(a.view(torch.int32), b.view(torch.int32)), (wire, tuple(p_.shape), float((a - b).abs().max()))
    if wire == "bf16":   # the 4 MiB weight sits in a bucket above the threshold: it really was rounded to the wire type
        big = max(params, key=lambda p_: p_.numel())
        assert torch.equal(flat(big.grad), flat(big.grad).to(torch.bfloat16).float())
    print("wire", wire, "bucket bytes", sizes, flush=True)
    # repeated use: optimizer steps through the hooked model
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    first = None
    for _ in range(5):
        opt.zero_grad(set_to_none=True)
        loss = nn.functional.cross_entropy(model(x), y)
        loss.backward()
        opt.step()
        first = float(loss.detach()) if first is None else first
    torch.cuda.synchronize()
    state.comm.check()
    last = float(loss.detach())
    assert last == last and last <= first, (first, last)   # finite, and SGD on one batch does not go up
    state.comm.destroy()
    del model, opt
print("DDP_HOOK_OK")
dist.destroy_process_group()
"""


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.gpu
def test_hook_inside_ddp_matches_stock_reducer_arithmetic():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    env = dict(os.environ, B200_TEST_ROOT=ROOT, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()),
               B200COLL_TIMEOUT_MS="20000")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    r = subprocess.run([sys.executable, "-c", SCRIPT], env=env, capture_output=True, text=True, timeout=170)
    assert r.returncode == 0 and "DDP_HOOK_OK" in r.stdout, r.stdout[-2000:] + "\n" + r.stderr[-4000:]
