"""Parity checks of the NRSfM stages shared by the CPU tier (emulated kernel sources) and the GPU
tier (CUDA library through the C ABI): `api` is the implementation under test, `orc` the oracle."""
import copy

import numpy as np

from defslam_b200 import nrsfm

# tolerances (fp64 path; the north star allows 1e-4 relative)
X_TOL = 1e-10       # Schwarp control points, absolute (values are O(1))
F32_TOL = 2e-6      # fp32 DiffProp records, relative to the record's scale
K_TOL = 1e-8        # normals (k1, k2), well-conditioned points
CTRL_TOL = 1e-8     # SfN control depths (median-normalised, O(1))


def identity_grid(bbs):
    NC = bbs.nptsu * bbs.nptsv
    X = np.zeros(2 * NC)
    k = 0
    for i in range(bbs.nptsu):
        for j in range(bbs.nptsv):
            X[k] = (bbs.umax - bbs.umin) * i / (bbs.nptsu - 1) + bbs.umin
            X[NC + k] = (bbs.vmax - bbs.vmin) * j / (bbs.nptsv - 1) + bbs.vmin
            k += 1
    return X


def check_schwarp_evaluate(api, orc, case, x):
    r0, J0 = orc.schwarp_evaluate(case, x)
    r1, J1 = api.schwarp_evaluate(case, x)
    assert np.abs(r0 - r1).max() <= 1e-12 * max(1.0, np.abs(r0).max())
    assert np.abs(J0 - J1).max() <= 1e-12 * np.abs(J0).max()


def check_diffprop(fa, fo):
    assert (fa.keep == fo.keep).all()
    for name in ("warp_uv", "J12", "J21", "H12"):
        a, o = getattr(fa, name), getattr(fo, name)
        assert np.abs(a - o).max() <= F32_TOL * max(1.0, np.abs(o).max()), name


def check_schwarp_fit(api, orc, case):
    fo = orc.schwarp_fit(case)
    fa = api.schwarp_fit(case)
    assert fa.d.iterations == fo.d.iterations
    assert fa.d.accepted == fo.d.accepted
    assert abs(fa.d.cost_initial - fo.d.cost_initial) <= 1e-9 * fo.d.cost_initial
    assert abs(fa.d.cost_final - fo.d.cost_final) <= 1e-9 * fo.d.cost_final
    assert np.abs(fa.x - fo.x).max() <= X_TOL
    check_diffprop(fa, fo)
    return fa, fo


def accepted_steps_case(case, iters=8):
    """Start from the identity warp so that the trust-region steps are accepted (from
    Warp::initialize the quirk-C6 Jacobian makes Ceres reject all three)."""
    c2 = copy.copy(case)
    c2.initialize = 0
    c2.x0 = identity_grid(case.bbs)
    c2.max_iterations = iters
    return c2


def check_normals(api, orc, ncase):
    no = orc.normals(ncase)
    na = api.normals(ncase)
    n = ncase.n
    assert (na.status[:n] == no.status[:n]).all()
    assert (na.pair_valid[:ncase.npairs] == no.pair_valid[:ncase.npairs]).all()
    ok = no.status[:n] == 1
    # conditioning: the covariance (J'J)^-1 says how far rounding may move the minimiser
    well = ok & (np.abs(no.cov).max(1) < 1e6) & (no.iters[:n] < ncase.max_iterations)
    assert well.sum() > 0.8 * ok.sum()
    same = na.iters[:n] == no.iters[:n]
    assert same[well].mean() > 0.99, same[well].mean()
    # the trust-region loop stops on a 1e-10 relative cost change: where both sides stop at the same
    # iteration the minimisers agree to rounding; a point that stops one iteration apart still agrees
    # far inside the north-star tolerance (1e-4)
    ws = well & same
    dk = np.abs(na.k - no.k).max(1)
    assert dk[ws].max() <= K_TOL, dk[ws].max()
    assert dk[ok].max() <= 1e-4, dk[ok].max()
    assert np.abs(na.normal - no.normal)[ws].max() <= 1e-6
    rel = np.abs(na.cov - no.cov)[ws] / np.abs(no.cov[ws]).max(1, keepdims=True)
    assert rel.max() <= 1e-5, rel.max()
    pv = no.pair_valid[:ncase.npairs] == 1
    pw = pv & np.repeat(ws, np.diff(ncase.pair_ptr))
    assert np.abs(na.pair_normal[:ncase.npairs] - no.pair_normal[:ncase.npairs])[pw].max() <= 1e-6
    return na, no


def check_sfn(api, orc, scase):
    A0, b0 = orc.sfn_system(scase)
    A1, b1 = api.sfn_system(scase)
    assert np.abs(A0 - A1).max() <= 1e-12 * np.abs(A0).max()
    assert np.abs(b0 - b1).max() == 0.0
    co, xo = orc.sfn_solve(scase)
    co, xo = co.copy(), xo.copy()
    ca, xa = api.sfn_solve(scase)
    assert np.abs(ca - co).max() <= CTRL_TOL
    assert np.abs(xa - xo).max() <= 1e-6
    return ca, co


def window_chain(api, orc, seed=1, n_keypoints=400, n_views=3, nptsu=nrsfm.NCU, nptsv=nrsfm.NCV):
    """One keyframe window through the three stages with `api`, every stage checked against the
    oracle fed with the same inputs."""
    win = nrsfm.make_window(seed, n_keypoints=n_keypoints, n_views=n_views, nptsu=nptsu, nptsv=nptsv)
    cases = nrsfm.schwarp_cases(win)
    fits = []
    for c in cases:
        fa, fo = check_schwarp_fit(api, orc, c)
        fits.append(fo)
    ncase = nrsfm.normals_case(win, fits)
    na, no = check_normals(api, orc, ncase)
    scase = nrsfm.sfn_case(win, no)
    check_sfn(api, orc, scase)
    return win


def check_sim3(api, orc, cases):
    ro = orc.sim3_register(cases)
    ra = api.sim3_register(cases)
    for a, o in zip(ra, ro):
        # the oracle differentiates numerically (delta 1e-9) like the reference, the kernel analytically:
        # the first run's trajectory is the same, the estimates agree to the noise of those differences
        assert a["iterations"][0] == o["iterations"][0]
        assert np.abs(a["rot"] - o["rot"]).max() < 1e-7 and np.abs(a["trans"] - o["trans"]).max() < 1e-7
        assert abs(a["scale"] - o["scale"]) < 1e-7 * o["scale"]
        assert abs(a["chi2"] - o["chi2"]) < 1e-6 * o["chi2"]
        assert a["inliers"] == o["inliers"] and a["acceptable"] == o["acceptable"]
    return ra, ro
