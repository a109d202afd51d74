"""Shared helpers for the GPU parity tests: seeded inputs and dtype tables."""
import torch

from ant_ray_b200 import _native as N

INT_DTYPES = [torch.int8, torch.uint8, torch.int32, torch.int64]
FLOAT_DTYPES = [torch.float16, torch.bfloat16, torch.float32, torch.float64]
NATIVE = {torch.int8: N.INT8, torch.uint8: N.UINT8, torch.int32: N.INT32, torch.int64: N.INT64,
          torch.float16: N.FLOAT16, torch.bfloat16: N.BFLOAT16, torch.float32: N.FLOAT32, torch.float64: N.FLOAT64}


def make_input(dtype, n, rank, op="sum"):
    """rank r draws from manual_seed(1234 + r); N(0,1) floats, bounded ints."""
    g = torch.Generator().manual_seed(1234 + rank)
    if dtype.is_floating_point:
        x = torch.randn(n, generator=g, dtype=torch.float32)
        if op == "prod":
            x = 1.0 + 0.1 * x  # keep W-fold products finite in half precision
        return x.to(dtype)
    info = torch.iinfo(dtype)
    lo, hi = max(info.min, -(2**15)), min(info.max, 2**15)
    if op == "prod":
        lo, hi = max(info.min, -3), min(info.max, 4)
    return torch.randint(lo, hi, (n,), generator=g, dtype=torch.int64).to(dtype)


def assert_equal_bits(a: torch.Tensor, b: torch.Tensor, what=""):
    """Bit-exact comparison (NaN-safe: compares the raw bytes)."""
    a, b = a.detach().cpu().contiguous(), b.detach().cpu().contiguous()
    assert a.dtype == b.dtype and a.shape == b.shape, f"{what}: {a.dtype}{tuple(a.shape)} vs {b.dtype}{tuple(b.shape)}"
    if not torch.equal(a.view(torch.uint8), b.view(torch.uint8)):
        diff = (a.double() - b.double()).abs()
        idx = int(diff.argmax())
        raise AssertionError(f"{what}: {int((diff > 0).sum())}/{a.numel()} elements differ; worst at {idx}: "
                             f"{a.flatten()[idx].item()} vs {b.flatten()[idx].item()}")
