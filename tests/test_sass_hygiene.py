"""The reduce kernels must compile without local memory: `cuobjdump -res-usage` of the built library.

No GPU needed (the listing is static); skipped when the CUDA binary utilities are not installed."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ant-ray_b200", "libb200coll.so")

# kernels allowed to keep one spilled scalar (8 bytes) outside their inner loops: profiles/r02_sass_local_memory.txt
ALLOWED = re.compile(r"k_allreduce_nvls_lanes|k_allreduce_twoshotI(dd|mm|ll)Li\dELi8E")


@pytest.fixture(scope="module")
def res_usage():
    if shutil.which("cuobjdump") is None or shutil.which("c++filt") is None:
        pytest.skip("cuobjdump / c++filt not installed")
    if not os.path.exists(LIB):
        pytest.skip("libb200coll.so not built")
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], check=True, capture_output=True, text=True).stdout
    usage, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "STACK:" in line:
            usage[cur] = (int(re.search(r"REG:(\d+)", line).group(1)), int(re.search(r"STACK:(\d+)", line).group(1)))
            cur = None
    assert len(usage) > 100, "res-usage listing not parsed"
    return usage


def test_no_kernel_uses_a_stack_except_the_listed_ones(res_usage):
    bad = {k: v for k, v in res_usage.items() if v[1] > 0 and not ALLOWED.search(k)}
    assert not bad, f"kernels with local memory: {bad}"
    assert all(v[1] <= 8 for v in res_usage.values())


def test_reduce_kernels_exist_per_world_size(res_usage):
    # one kernel per world size 2 / 4 / 8 and one for the others (WT = 0): fp32 SUM two-shot
    for wt in (0, 2, 4, 8):
        assert any(re.search(rf"k_allreduce_twoshotIffLi0ELi{wt}E", k) for k in res_usage), wt
    # two CTAs of 512 threads per SM: at most 64 registers per thread in every reducing kernel
    for k, (reg, _) in res_usage.items():
        if re.search(r"k_allreduce_(oneshot|twoshot|nvls)|k_reduce|k_reducescatter", k):
            assert reg <= 64, (k, reg)
