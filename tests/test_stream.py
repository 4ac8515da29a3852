"""Tracking + mapping loop (config C3: SfT every frame, keyframes, NRSfM window -> new template):
the same loop through the kernel code and through the oracle, frame by frame."""
import ctypes as C

import numpy as np
import pytest

from defslam_b200 import _capi, sft, stream
from tests.helpers import emu_lib

SHORT = dict(G=13, n_points=500, n_frames=26, kf_every=4, nrsfm_every_kf=5, n_views=3)
NODE_TOL = 1e-4   # north_star: <= 1e-4 relative on node positions


def _oracle_backend(oracle):
    return stream.Backend(oracle.load(), "oracle_", oracle.sft_solve)


def _emu_backend():
    lib = emu_lib()

    def solve(frame):
        out = sft.SftOutput(frame.template.n_nodes, frame.n_matches)
        p = frame.problem()
        res = _capi.SftResult()
        out.fill(res)
        rc = lib.emu_sft_solve(C.byref(p), C.byref(res))
        assert rc == 0
        out.r = res
        return out
    lib.emu_sft_solve.restype = C.c_int
    lib.emu_sft_solve.argtypes = [C.POINTER(_capi.SftProblem), C.POINTER(_capi.SftResult)]
    return stream.Backend(lib, "emu_", solve)


def _compare(a, b):
    assert a.n_nrsfm == b.n_nrsfm == 1
    assert a.n_template_updates == b.n_template_updates == 1      # the NRSfM surface was accepted and swapped in
    assert len(a.nodes_cam) == len(b.nodes_cam)
    worst = 0.0
    for x, y in zip(a.nodes_cam, b.nodes_cam):
        worst = max(worst, np.abs(x - y).max() / np.sqrt((y ** 2).sum(1).mean()))
    assert worst <= NODE_TOL, worst
    assert np.allclose(a.rmse, b.rmse, atol=1e-4)
    assert np.allclose(a.template_rmse, b.template_rmse, atol=1e-4)
    return worst


def test_stream_kernel_code_matches_oracle(oracle):
    cfg = stream.StreamConfig(**SHORT)
    ro = stream.run_stream(_oracle_backend(oracle), cfg)
    re = stream.run_stream(_emu_backend(), cfg)
    _compare(re, ro)
    assert max(ro.rmse) < 0.08 and min(ro.inliers) > 0.9 * 500 * 0.9


@pytest.mark.gpu
def test_stream_gpu_matches_oracle(cuda_lib, oracle):
    cfg = stream.StreamConfig(**SHORT)
    ro = stream.run_stream(_oracle_backend(oracle), cfg)
    rg = stream.run_stream(stream.cuda_backend(), cfg)
    _compare(rg, ro)


@pytest.mark.gpu
def test_stream_c3_shape_runs(cuda_lib):
    """the C3 shape itself (17 x 17 mesh, 600 points, 17 x 17 control grid), first NRSfM event included"""
    cfg = stream.StreamConfig(n_frames=60, n_views=4)
    r = stream.run_stream(stream.cuda_backend(), cfg, keep_nodes=False)
    assert r.n_nrsfm == 1 and len(r.rmse) == 60
    assert max(r.rmse) < 0.1
