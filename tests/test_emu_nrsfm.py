"""CPU tier: the NRSfM kernel sources compiled with g++ as a one-thread team (tests/emu/) against
the oracle.  Checks the kernels' own arithmetic, gather indexing, block-banded factorisation and LM
control flow without a GPU."""
import copy
import ctypes as C

import numpy as np
import pytest

from defslam_b200 import nrsfm
from tests import helpers
from tests import nrsfm_checks as ck


@pytest.fixture(scope="module")
def apis(oracle):
    from tests.emu import build
    return nrsfm.Api(C.CDLL(build.build()), "emu_"), nrsfm.Api(oracle.load(), "oracle_")


def test_schwarp_evaluate(apis):
    api, orc = apis
    win = nrsfm.make_window(3, n_keypoints=200, n_views=1)
    c = nrsfm.schwarp_cases(win)[0]
    ck.check_schwarp_evaluate(api, orc, c, orc.schwarp_init(c))


@pytest.mark.parametrize("grid", [(13, 15), (9, 9), (7, 12)])
def test_window_chain(apis, grid):
    api, orc = apis
    ck.window_chain(api, orc, seed=5, n_keypoints=300, n_views=2, nptsu=grid[0], nptsv=grid[1])


def test_schwarp_accepted_steps(apis):
    api, orc = apis
    win = nrsfm.make_window(7, n_keypoints=300, n_views=1)
    c = ck.accepted_steps_case(nrsfm.schwarp_cases(win)[0])
    fa, fo = ck.check_schwarp_fit(api, orc, c)
    assert fo.d.accepted >= 1


def test_schwarp_out_of_domain(apis):
    api, _ = apis
    win = nrsfm.make_window(2, n_keypoints=100, n_views=1)
    c = nrsfm.schwarp_cases(win)[0]
    c.kp1 = c.kp1.copy()
    c.kp1[3, 0] = c.bbs.umax + 0.5
    with pytest.raises(nrsfm.DefslamError) as e:
        api.schwarp_fit(c)
    assert e.value.rc == -1


def test_normals_mixed_pairs(apis):
    api, orc = apis
    win = nrsfm.make_window(13, n_keypoints=500, n_views=4, match_frac=0.6)
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    nc = nrsfm.normals_case(win, fits)
    rng = np.random.default_rng(0)
    nc.pair_from_ref = (rng.uniform(size=nc.npairs) > 0.2).astype(np.uint8)
    nc.k_first = rng.normal(size=(nc.npairs, 2)).astype(np.float32) * 0.1
    nc.k_first[rng.uniform(size=nc.npairs) > 0.7] = np.nan
    ck.check_normals(api, orc, nc)
    nc2 = copy.copy(nc)
    nc2.corrected_t2 = 1
    ck.check_normals(api, orc, nc2)


def test_sim3_registration_and_min_median_scale(apis):
    api, orc = apis
    cases = [nrsfm.sim3_case(s) for s in range(3)] + [nrsfm.sim3_case(9, n=40, noise=1e-4, outlier_frac=0.0)]
    ra, ro = ck.check_sim3(api, orc, cases)
    assert ro[-1]["acceptable"] == 1 and ro[0]["acceptable"] == 0
    for c in cases:
        assert api.scale_min_median(c.pts1, c.pts2, seed=7) == orc.scale_min_median(c.pts1, c.pts2, seed=7)
