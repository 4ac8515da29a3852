"""GPU parity tests of the NRSfM stages: CUDA path (through the C ABI) vs the CPU oracle."""
import copy

import numpy as np
import pytest

from defslam_b200 import nrsfm
from tests import nrsfm_checks as ck

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def apis(oracle, cuda_lib):
    return nrsfm.Api(cuda_lib, "defslam_"), nrsfm.Api(oracle.load(), "oracle_")


def test_schwarp_evaluate_matches_oracle(apis):
    api, orc = apis
    win = nrsfm.make_window(3, n_keypoints=300, n_views=1)
    c = nrsfm.schwarp_cases(win)[0]
    x = orc.schwarp_init(c)
    ck.check_schwarp_evaluate(api, orc, c, x)
    ck.check_schwarp_evaluate(api, orc, c, ck.identity_grid(c.bbs))


@pytest.mark.parametrize("grid", [(13, 15), (9, 9), (17, 17)])
def test_window_chain_matches_oracle(apis, grid):
    api, orc = apis
    ck.window_chain(api, orc, seed=5, n_keypoints=500, n_views=2, nptsu=grid[0], nptsv=grid[1])


def test_schwarp_accepted_steps(apis):
    api, orc = apis
    win = nrsfm.make_window(7, n_keypoints=600, n_views=1)
    c = ck.accepted_steps_case(nrsfm.schwarp_cases(win)[0])
    fa, fo = ck.check_schwarp_fit(api, orc, c)
    assert fo.d.accepted >= 1


def test_schwarp_batched_equals_single(apis):
    api, orc = apis
    win = nrsfm.make_window(11, n_keypoints=800, n_views=5, match_frac=0.4)
    cases = nrsfm.schwarp_cases(win)
    cases[2] = ck.accepted_steps_case(cases[2])
    outs = api.schwarp_fit_batched(cases * 40)  # 200 pairs > 148 CTAs: exercises the work counter
    for i, c in enumerate(cases):
        single = api.schwarp_fit(c)
        for rep in (i, i + 5 * 39):
            assert np.array_equal(outs[rep].x, single.x)
            assert np.array_equal(outs[rep].J12, single.J12)
            assert outs[rep].d.iterations == single.d.iterations


def test_schwarp_pipelined_batch_equals_single(apis):
    """batches of >= 2 waves run as a pipeline of chunks over three streams (upload / kernels / download): same results,
    fit for fit, as single calls; a fit that fails (site outside the domain) leaves its outputs untouched"""
    api, orc = apis
    win = nrsfm.make_window(12, n_keypoints=500, n_views=5, match_frac=0.5)
    cases = nrsfm.schwarp_cases(win)
    cases[1] = ck.accepted_steps_case(cases[1])
    singles = [api.schwarp_fit(c) for c in cases]
    reps = 70
    batch = cases * reps  # 350 pairs: three chunks on a 148-SM device
    bad = copy.copy(batch[177])
    bad.kp1 = bad.kp1.copy()
    bad.kp1[3, 0] = bad.bbs.umax + 0.5
    batch[177] = bad
    with pytest.raises(nrsfm.DefslamError):
        api.schwarp_fit_batched(batch)  # the failed fit is reported; the others are complete
    batch[177] = cases[177 % len(cases)]
    outs = api.schwarp_fit_batched(batch)
    for rep in (0, 147, 148, 177, 296, len(batch) - 1):
        single = singles[rep % len(cases)]
        assert np.array_equal(outs[rep].x, single.x)
        assert np.array_equal(outs[rep].J12, single.J12) and np.array_equal(outs[rep].H12, single.H12)
        assert np.array_equal(outs[rep].keep, single.keep)
        assert outs[rep].d.iterations == single.d.iterations and outs[rep].d.accepted == single.d.accepted


def test_schwarp_out_of_domain_is_rejected(apis):
    api, _ = apis
    win = nrsfm.make_window(2, n_keypoints=200, n_views=1)
    c = nrsfm.schwarp_cases(win)[0]
    c.kp1 = c.kp1.copy()
    c.kp1[3, 0] = c.bbs.umax + 0.5
    with pytest.raises(nrsfm.DefslamError) as e:
        api.schwarp_fit(c)
    assert e.value.rc == -1


def test_normals_large_batch_and_mixed_pairs(apis):
    api, orc = apis
    win = nrsfm.make_window(13, n_keypoints=3000, n_views=6, match_frac=0.6)
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    nc = nrsfm.normals_case(win, fits)
    # some pairs do not start at the reference keyframe: they only receive a transferred normal
    rng = np.random.default_rng(0)
    nc.pair_from_ref = (rng.uniform(size=nc.npairs) > 0.2).astype(np.uint8)
    nc.k_first = rng.normal(size=(nc.npairs, 2)).astype(np.float32) * 0.1
    nc.k_first[rng.uniform(size=nc.npairs) > 0.7] = np.nan
    ck.check_normals(api, orc, nc)
    nc2 = copy.copy(nc)
    nc2.corrected_t2 = 1
    ck.check_normals(api, orc, nc2)


def test_normals_pipelined_batch_equals_single_pass(apis):
    """batches of >= 64 k map points run as a pipeline of chunks over two streams: same results, point for point,
    as the tile they are made of solved in one pass (untouched outputs stay untouched)"""
    api, orc = apis
    win = nrsfm.make_window(21, n_keypoints=1500, n_views=4, match_frac=0.8)
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    nc = nrsfm.normals_case(win, fits)
    rng = np.random.default_rng(1)
    nc.pair_from_ref = (rng.uniform(size=nc.npairs) > 0.2).astype(np.uint8)
    nc.k_first = rng.normal(size=(nc.npairs, 2)).astype(np.float32) * 0.1
    one = api.normals(nc)
    reps = 70000 // nc.n + 1
    ptr = [0]
    for _ in range(reps):
        ptr.extend((nc.pair_ptr[1:] + ptr[-1]).tolist())
    tile = lambda a: np.ascontiguousarray(np.concatenate([a] * reps))
    big = nrsfm.NormalsCase(pair_ptr=np.array(ptr, np.int32), J12=tile(nc.J12[:nc.npairs]), J21=tile(nc.J21[:nc.npairs]),
                            H12=tile(nc.H12[:nc.npairs]), I1=tile(nc.I1[:nc.npairs]), I2=tile(nc.I2[:nc.npairs]),
                            pair_from_ref=tile(nc.pair_from_ref[:nc.npairs]), k_first=tile(nc.k_first[:nc.npairs]),
                            k_init=tile(nc.k_init), ref_uv=tile(nc.ref_uv))
    assert big.n >= 65536
    out = api.normals(big)
    n, q = nc.n, nc.npairs
    for r in (0, 1, reps // 2, reps - 1):
        assert np.array_equal(out.status[r * n:(r + 1) * n], one.status[:n])
        assert np.array_equal(out.iters[r * n:(r + 1) * n], one.iters[:n])
        assert np.array_equal(out.k[r * n:(r + 1) * n], one.k, equal_nan=True)
        assert np.array_equal(out.cov[r * n:(r + 1) * n], one.cov, equal_nan=True)
        assert np.array_equal(out.normal[r * n:(r + 1) * n], one.normal, equal_nan=True)
        assert np.array_equal(out.pair_valid[r * q:(r + 1) * q], one.pair_valid[:q])
        assert np.array_equal(out.pair_normal[r * q:(r + 1) * q], one.pair_normal[:q], equal_nan=True)


def test_polysolver_coefficients_match_oracle(apis):
    api, orc = apis
    rng = np.random.default_rng(1)
    n = 4096
    J12 = (np.eye(2).reshape(1, 4) + rng.normal(size=(n, 4)) * 0.2).astype(np.float32)
    H12 = (rng.normal(size=(n, 6)) * 0.3).astype(np.float32)
    I1 = rng.uniform(-0.7, 0.7, (n, 2)).astype(np.float32)
    I2 = rng.uniform(-0.7, 0.7, (n, 2)).astype(np.float32)
    a1, a2 = api.polysolver_coefficients(J12, H12, I1, I2)
    o1, o2 = orc.polysolver_coefficients(J12, H12, I1, I2)
    s = np.maximum(np.abs(o1).max(1), np.abs(o2).max(1))[:, None]
    assert (np.abs(a1 - o1) / s).max() < 1e-13 and (np.abs(a2 - o2) / s).max() < 1e-13


def test_sfn_batched_and_few_normals(apis):
    api, orc = apis
    cases = []
    for seed in range(3):
        win = nrsfm.make_window(20 + seed, n_keypoints=300, n_views=2)
        fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
        no = orc.normals(nrsfm.normals_case(win, fits))
        cases.append(nrsfm.sfn_case(win, no))
    few = copy.copy(cases[0])
    few.uv, few.normals = few.uv[:10].copy(), few.normals[:10].copy()
    cases.append(few)
    refs = []
    for c in cases:
        co, xo = orc.sfn_solve(c)
        refs.append((co.copy(), xo.copy()))
    rcs = api.sfn_solve_batched(cases)
    assert (rcs == 0).all()
    for c, (co, xo) in zip(cases, refs):
        assert np.abs(c.ctrl - co).max() <= ck.CTRL_TOL
        assert np.abs(c.xyz - xo).max() <= 1e-6


def test_sfn_packed_fallbacks_agree_with_the_tile_path(apis):
    """N as tiles on the tensor cores (default), packed in shared memory (mode 1), packed in the workspace (mode 0: the
    path of grids too large for shared memory): all three against the QR oracle, and against each other"""
    import os
    api, orc = apis
    win = nrsfm.make_window(31, n_keypoints=400, n_views=2)
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    no = orc.normals(nrsfm.normals_case(win, fits))
    base = nrsfm.sfn_case(win, no)
    co, xo = orc.sfn_solve(copy.copy(base))
    co, xo = co.copy(), xo.copy()
    got = {}
    for mode in ("2", "1", "0"):
        os.environ["DEFSLAM_SFN_MODE"] = mode
        try:
            c = copy.copy(base)
            c.ctrl, c.xyz = None, None  # fresh output arrays per mode
            assert (api.sfn_solve_batched([c]) == 0).all()
        finally:
            del os.environ["DEFSLAM_SFN_MODE"]
        assert np.abs(c.ctrl - co).max() <= ck.CTRL_TOL and np.abs(c.xyz - xo).max() <= 1e-6
        got[mode] = c.ctrl.copy()
    assert np.abs(got["2"] - got["1"]).max() <= 1e-9 and np.array_equal(got["1"], got["0"])


def test_sfn_nan_normals_fail_loudly(apis):
    api, orc = apis
    win = nrsfm.make_window(4, n_keypoints=200, n_views=1)
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    no = orc.normals(nrsfm.normals_case(win, fits))
    c = nrsfm.sfn_case(win, no)
    c.normals = c.normals.copy()
    c.normals[0] = np.nan
    with pytest.raises(nrsfm.DefslamError) as e:
        api.sfn_solve(c)
    assert e.value.rc == -3


def test_sim3_registration_and_min_median_scale(apis):
    api, orc = apis
    cases = [nrsfm.sim3_case(s, n=1200) for s in range(6)] + [nrsfm.sim3_case(9, n=40, noise=1e-4, outlier_frac=0.0)]
    ra, ro = ck.check_sim3(api, orc, cases * 30)       # 210 keyframes in one launch
    assert ro[6]["acceptable"] == 1 and ro[0]["acceptable"] == 0
    for c in cases[:3]:
        a, o = api.scale_min_median(c.pts1, c.pts2, seed=7), orc.scale_min_median(c.pts1, c.pts2, seed=7)
        assert abs(a - o) <= 1e-6 * abs(o)


def test_schwarp_initial_matches_oracle(apis):
    """DefORBmatcher::CalculateInitialSchwarp (DefORBmatcher.cc:111-187): Warp::initialize, the loss-corrected
    residuals of the Warp block and the > 20 filter, CUDA vs oracle; a few wrong matches so that the filter acts"""
    api, orc = apis
    for seed in (5, 6):
        w = nrsfm.make_window(seed, n_keypoints=400, n_views=2)
        c = nrsfm.schwarp_cases(w)[0]
        c.kp2 = c.kp2.copy()
        c.kp2[:8] += 0.3
        xa, ka, ea = api.schwarp_initial(c)
        xo, ko, eo = orc.schwarp_initial(c)
        assert np.abs(xa - xo).max() < 1e-9
        assert np.allclose(ea, eo, rtol=1e-8, atol=1e-12)
        assert np.array_equal(ka, ko) and 0 < ka.sum() < c.n
