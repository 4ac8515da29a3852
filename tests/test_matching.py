"""Projection search of template points (DefORBmatcher::SearchByProjection,
Modules/Matching/DefORBmatcher.cc:296-451): the parallel candidates + in-order resolve of the kernel
against the oracle's literal loop (grid of keypoint lists, first minimum, hidden keypoints)."""
import numpy as np
import pytest

from defslam_b200 import matching
from tests.helpers import emu_lib

CASES = [dict(seed=1), dict(seed=2, n_last=300, n_clutter=2000), dict(seed=3, stereo=True),
         dict(seed=4, n_last=1500, move_px=8.0, flip_bits=40)]


def _check(lib, prefix, oracle):
    olib = oracle.load()
    for kw in CASES + [dict(seed=6, n_last=250, n_clutter=5000, th=45.0)]:   # last: > 64 candidates per point
        for orient in (1, 0):
            kw = dict(kw)
            th = kw.pop("th", None)
            c = matching.make_case(**kw)
            if th is not None:
                c.th = th
            c.check_orientation = orient
            mo, no = matching.search_by_projection(c, olib, "oracle_")
            m, n = matching.search_by_projection(c, lib, prefix)
            assert n == no
            assert np.array_equal(m, mo)                     # index work: bit-exact
            good = mo >= 0
            assert good.sum() > 0.3 * (c.truth >= 0).sum()
            assert (mo[good] == c.truth[good]).mean() > 0.97  # and the matches are the right ones


def test_kernel_code_matches_oracle(oracle):
    _check(emu_lib(), "emu_", oracle)


def test_order_dependence_is_reproduced(oracle):
    """two map points whose best keypoint is the same: the earlier one takes it, the later one falls
    back to its second choice (or none) -- exactly like the reference's sequential loop"""
    c = matching.make_case(seed=7, n_last=40, n_clutter=0, flip_bits=1, move_px=0.2)
    c.last_state[:] = 1; c.last_has_obs[:] = 1; c.cur_taken[:] = 0; c.check_orientation = 0
    # map point 1 becomes a copy of map point 0 (same position, same descriptor)
    c.last_world_xyz[1] = c.last_world_xyz[0]; c.last_desc[1] = c.last_desc[0]; c.last_octave[1] = c.last_octave[0]
    mo, no = matching.search_by_projection(c, oracle.load(), "oracle_")
    me, ne = matching.search_by_projection(c, emu_lib(), "emu_")
    assert np.array_equal(mo, me) and no == ne
    assert (mo == 0).sum() == 1 and (mo == 1).sum() <= 1
    j0 = int(np.flatnonzero(mo == 0)[0])
    assert c.truth[j0] in (0, 1)


def test_empty_and_degenerate(oracle):
    c = matching.make_case(seed=5, n_last=50, n_clutter=10)
    c.last_state[:] = 0
    for lib, pre in ((oracle.load(), "oracle_"), (emu_lib(), "emu_")):
        m, n = matching.search_by_projection(c, lib, pre)
        assert n == 0 and (m == -1).all()


@pytest.mark.gpu
def test_gpu_matches_oracle(cuda_lib, oracle):
    _check(cuda_lib, "defslam_", oracle)
    c = matching.make_case(seed=5, n_last=50, n_clutter=10)
    c.last_state[:] = 0
    m, n = matching.search_by_projection(c)
    assert n == 0 and (m == -1).all()


# ----------------------------------------------------------------------------- warp-guided search
def _check_warp(lib, prefix, oracle):
    olib = oracle.load()
    for seed, kw in ((1, {}), (2, dict(n1=300, n_clutter=3000)), (3, dict(nptsu=17, nptsv=17))):
        c = matching.make_warp_case(seed, **kw)
        mo, no = matching.search_by_schwarp(c, olib, "oracle_")
        m, n = matching.search_by_schwarp(c, lib, prefix)
        assert n == no and np.array_equal(m, mo)
        good = mo >= 0
        assert good.sum() > 0.3 * c.kp1_state.sum()
        assert (c.truth[mo[good]] == np.flatnonzero(good)).mean() > 0.97
        assert not c.kp2_has_mp[mo[good]].any() and c.kp1_state[good].all()


def test_warp_search_kernel_code_matches_oracle(oracle):
    _check_warp(emu_lib(), "emu_", oracle)


@pytest.mark.gpu
def test_warp_search_gpu_matches_oracle(cuda_lib, oracle):
    _check_warp(cuda_lib, "defslam_", oracle)
    c = matching.make_warp_case(4, n1=20, n_clutter=5)
    c.kp1_state[:] = 0
    m, n = matching.search_by_schwarp(c)
    assert n == 0 and (m == -1).all()
