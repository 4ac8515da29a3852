"""New map points / exploration test (DefLocalMapping::CreateNewMapPoints, ::needNewTemplate,
Modules/Mapping/DefLocalMapping.cc:240-347,355-403): oracle and kernel arithmetic against vectors
produced by OpenCV itself (tests/golden/newpts_cv.npz, made by tests/golden/make_golden_cv.py)."""
import ctypes as C
import os

import numpy as np
import pytest

from defslam_b200 import _capi, nrsfm
from tests.helpers import emu_lib

GOLD = os.path.join(os.path.dirname(__file__), "golden", "newpts_cv.npz")


def _cases():
    g = np.load(GOLD)
    for ci, (seed, n, rows, cols, bh) in enumerate(g["cases"]):
        yield ci, int(rows), int(cols), {k: g[f"{k}{ci}"] for k in ("xy", "state", "surf", "T", "action", "world", "nnew")}


def _check(api, placement=True):
    for ci, rows, cols, c in _cases():
        act, world, n_new = api.new_map_points(c["xy"], c["state"], rows, cols, c["surf"], c["T"])
        assert np.array_equal(act, c["action"]), f"case {ci}"
        assert n_new == int(c["nnew"])
        assert np.array_equal(world, c["world"]), f"case {ci}"      # fp32 bit-exact vs cv2.gemm
        act2, world2, n_new2 = api.new_map_points(c["xy"], c["state"], rows, cols)   # count only
        assert world2 is None and n_new2 == n_new and np.array_equal(act2, act)


def test_oracle_matches_opencv(oracle):
    _check(nrsfm.Api(oracle.load(), "oracle_"))


def test_kernel_arithmetic_matches_opencv():
    _check(nrsfm.Api(emu_lib(), "emu_"))


def test_window_predicate_exhaustive(oracle):
    """every candidate pixel of a small image against a literal filter2D evaluation (the oracle)"""
    rows, cols = 23, 41  # kernel edge 2
    rng = np.random.default_rng(3)
    marked = np.stack([rng.integers(0, cols, 6), rng.integers(0, rows, 6)], 1)
    yy, xx = np.mgrid[0:rows, 0:cols]
    cand = np.stack([xx.ravel(), yy.ravel()], 1)
    xy = np.concatenate([marked, cand]).astype(np.float32) + np.float32(0.25)
    state = np.concatenate([np.ones(6, np.uint8), np.zeros(len(cand), np.uint8)])
    a = nrsfm.Api(oracle.load(), "oracle_").new_map_points(xy, state, rows, cols)
    b = nrsfm.Api(emu_lib(), "emu_").new_map_points(xy, state, rows, cols)
    assert np.array_equal(a[0], b[0]) and a[2] == b[2]
    assert 0 < a[2] < len(cand)


@pytest.mark.gpu
def test_gpu_matches_opencv_and_oracle(cuda_lib, oracle):
    api = nrsfm.Api()
    _check(api)
    # a frame-sized random case against the oracle, and the argument checks
    rng = np.random.default_rng(5)
    n, rows, cols = 5000, 480, 640
    xy = np.stack([rng.uniform(0, cols - 0.01, n), rng.uniform(0, rows - 0.01, n)], 1).astype(np.float32)
    state = rng.choice([0, 1, 2], n, p=[0.9, 0.06, 0.04]).astype(np.uint8)
    surf = rng.normal(size=(n, 3)).astype(np.float32)
    T = np.eye(4, dtype=np.float32); T[:3, 3] = [0.1, -0.2, 0.05]
    a = api.new_map_points(xy, state, rows, cols, surf, T)
    b = nrsfm.Api(oracle.load(), "oracle_").new_map_points(xy, state, rows, cols, surf, T)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2] == b[2]
    assert 0 < a[2] < n
    xy[7] = [cols + 3, 5]
    with pytest.raises(nrsfm.DefslamError) as e:
        api.new_map_points(xy, state, rows, cols)
    assert e.value.rc == -1
    assert api.new_map_points(xy[:0], state[:0], rows, cols)[2] == 0
