"""Tensor-header ring (N3): binary (shape, dtype) records next to the cell ring instead of a pickled
message on a second channel (reference torch_tensor_accelerator_channel.py:574-578, 592-608)."""
import threading

import pytest
import torch

from mini_actor import get, spawn

from ant_ray_b200.header_ring import SLOTS, HeaderRing, HeaderTimeout, ring_path


def test_round_trip_and_order(tmp_path):
    path = str(tmp_path / "ring")
    w, r = HeaderRing(path, "w"), HeaderRing(path, "r")
    specs = [((3, 4), torch.float16), ((), torch.int64), ((7,), torch.bfloat16), ((2, 3, 4, 5, 6, 7, 8, 9), torch.float32), ((0, 5), torch.uint8)]
    for i, (shape, dtype) in enumerate(specs):
        w.put(shape, dtype, i, len(specs))
    for i, (shape, dtype) in enumerate(specs):
        assert r.get(1) == (shape, dtype, i, len(specs))
    with pytest.raises(HeaderTimeout):
        r.get(0.05)
    with pytest.raises(ValueError):
        w.put((1,) * 9, torch.float32)
    w.close(unlink=True)
    r.close(unlink=True)


def test_wraps_and_applies_back_pressure(tmp_path):
    path = str(tmp_path / "ring")
    w, r = HeaderRing(path, "w"), HeaderRing(path, "r")
    for i in range(SLOTS):
        w.put((i,), torch.float32)
    with pytest.raises(HeaderTimeout):
        w.put((99,), torch.float32, timeout_s=0.1)  # full: the reader has consumed nothing
    got = []
    t = threading.Thread(target=lambda: [got.append(r.get(10)[0][0]) for _ in range(3 * SLOTS)])
    t.start()
    for i in range(SLOTS, 3 * SLOTS):
        w.put((i,), torch.float32, timeout_s=10)
    t.join(20)
    assert got == list(range(3 * SLOTS))


def test_cancel_releases_a_blocked_reader(tmp_path):
    r = HeaderRing(str(tmp_path / "ring"), "r")
    flag = []
    threading.Timer(0.1, lambda: flag.append(1)).start()
    with pytest.raises(HeaderTimeout):
        r.get(30, cancelled=lambda: bool(flag))


class _Reader:
    def __init__(self, path):
        self.ring = HeaderRing(path, "r")

    def read(self, n):
        return [self.ring.get(20) for _ in range(n)]


def test_between_two_processes(tmp_path):
    path = ring_path("some/key", 0, 1).replace("/dev/shm", str(tmp_path))
    a = spawn(_Reader, path)
    try:
        ref = a.read.remote(200)
        w = HeaderRing(path, "w")
        for i in range(200):
            w.put((i, 2 * i), torch.float16 if i % 2 else torch.int32, i % 5, 5, timeout_s=20)
        res = get(ref, timeout=60)
        assert res == [((i, 2 * i), torch.float16 if i % 2 else torch.int32, i % 5, 5) for i in range(200)]
    finally:
        a.kill()
