"""Every product module imports on a CPU-only box (nothing CUDA-dependent at import time)."""
import importlib
import os
import pkgutil

import ant_ray_b200


def test_every_module_imports():
    names = [m.name for m in pkgutil.iter_modules(ant_ray_b200.__path__)]
    assert {"collective", "b200_group", "communicator", "channel", "ddp_hook", "train", "rendezvous", "loopback", "_native",
            "types", "experimental_collective", "rdt_transport", "collective_op", "channel_context"} <= set(names)
    for n in names:
        if n == "_ray_actors" or n.startswith("lib"):  # needs Ray / is the native library, not a Python module
            continue
        importlib.import_module(f"ant_ray_b200.{n}")


def test_package_layout():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for rel in ("include/b200coll.h", "ant-ray_b200/csrc/b200coll.cu", "ant-ray_b200/csrc/coll_kernels.cuh", "oracle/oracle_reduce.c",
                "DESIGN.md", "INTEGRATION.md", "bench.py", "__graft_entry__.py", "tests/golden/gloo_vectors.pt", "tests/golden/make_golden.py"):
        assert os.path.exists(os.path.join(root, rel)), rel
