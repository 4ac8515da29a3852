"""Pins of the NRSfM oracle (oracle/nrsfm_oracle.c) -- and, where noted, of the kernel sources too.

The reference ships no tests for this path and solves with Ceres/Eigen (absent here), so:
  * the polynomial coefficients are pinned to the REFERENCE's own PolySolver::getCoefficients
    (golden vectors from oracle/_ref/libpolysolver_ref.so, and live when the library is present);
  * exact Jacobians are pinned by finite differences; the deliberately inexact data Jacobian
    (quirk C6) by a transcription test;
  * the isometric polynomial system is pinned by substituting a rigidly moving plane (a homography),
    which also characterises quirk C7 numerically;
  * solutions are pinned by optimality conditions and by committed regression vectors."""
import copy
import ctypes as C
import os

import numpy as np
import pytest

from defslam_b200 import _capi, nrsfm
from tests import helpers
from tests.golden import make_golden as mg


@pytest.fixture(scope="module")
def orc(oracle):
    return nrsfm.Api(oracle.load(), "oracle_")


@pytest.fixture(scope="module")
def emu():
    from tests.emu import build
    return nrsfm.Api(C.CDLL(build.build()), "emu_")


@pytest.fixture(scope="module")
def win():
    return nrsfm.make_window(42, n_keypoints=300, n_views=2)


# --------------------------------------------------------------------------- polynomials ---
def test_polysolver_coefficients_match_reference_golden(orc, emu):
    g = helpers.golden("polysolver_ref.npz")
    for api in (orc, emu):
        e1, e2 = api.polysolver_coefficients(g["J12"], g["H12"], g["I1"], g["I2"])
        s = np.maximum(np.abs(g["eq1"]).max(1), np.abs(g["eq2"]).max(1))[:, None]
        assert (np.abs(e1 - g["eq1"]) / s).max() < 1e-14
        assert (np.abs(e2 - g["eq2"]) / s).max() < 1e-14


def test_polysolver_live_reference_when_present(orc):
    path = os.path.join(helpers.ROOT, "oracle", "_ref", "libpolysolver_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libpolysolver_ref.so not built")
    ref = C.CDLL(path)
    ref.ref_polysolver_coefficients.argtypes = [C.c_double] * 12 + [_capi.c_double_p] * 2
    J12, H12, I1, I2 = mg.poly_inputs(n=500, seed=99)
    t1, t2, e1, e2 = mg.poly_scalars(J12, H12, I1, I2)
    o1, o2 = orc.polysolver_coefficients(J12, H12, I1, I2)
    q1, q2 = np.zeros(10), np.zeros(10)
    for i in range(len(J12)):
        ref.ref_polysolver_coefficients(*J12[i], t1[i], t2[i], e1[i], e2[i], *I1[i], *I2[i],
                                        _capi.as_ptr(q1, C.c_double), _capi.as_ptr(q2, C.c_double))
        s = max(np.abs(q1).max(), np.abs(q2).max())
        assert np.abs(o1[i] - q1).max() <= 1e-14 * s and np.abs(o2[i] - q2).max() <= 1e-14 * s


def _plane_pair(seed):
    """A plane n.X = 1 seen before and after a rigid motion: the warp is the homography R + t n'."""
    rng = np.random.default_rng(seed)
    nrm = np.array([0.2, -0.1, 1.0]) + rng.normal(size=3) * 0.1
    nrm /= np.linalg.norm(nrm)
    R = nrsfm._rot(rng.normal(size=3), 0.15)
    t = rng.normal(size=3) * 0.1
    Hm = R + np.outer(t, nrm)

    def warp(q):
        p = Hm @ np.array([q[0], q[1], 1.0])
        return p[:2] / p[2]
    q1 = rng.uniform(-0.4, 0.4, 2)
    e = 1e-4
    f0 = warp(q1)
    fu1, fu0, fv1, fv0 = warp(q1 + [e, 0]), warp(q1 - [e, 0]), warp(q1 + [0, e]), warp(q1 - [0, e])
    du, dv = (fu1 - fu0) / (2 * e), (fv1 - fv0) / (2 * e)
    duu, dvv = (fu1 - 2 * f0 + fu0) / e ** 2, (fv1 - 2 * f0 + fv0) / e ** 2
    duv = (warp(q1 + [e, e]) - warp(q1 + [e, -e]) - warp(q1 + [-e, e]) + warp(q1 - [e, e])) / (4 * e * e)
    J12 = np.array([[du[0], du[1], dv[0], dv[1]]], np.float32)
    H12 = np.array([[duu[0], duu[1], duv[0], duv[1], dvv[0], dvv[1]]], np.float32)
    det = du[0] * dv[1] - dv[0] * du[1]
    J21 = np.array([[dv[1] / det, -dv[0] / det, -du[1] / det, du[0] / det]], np.float32)
    k1 = nrm[:2] / (nrm @ np.array([q1[0], q1[1], 1.0]))           # normal ~ (k1, k2, 1 - k1 u - k2 v)
    n2 = R @ nrm
    k2 = n2[:2] / (n2 @ np.array([f0[0], f0[1], 1.0]))
    return J12, J21, H12, q1.astype(np.float32)[None], f0.astype(np.float32)[None], k1, k2


def _normals_case(J12, J21, H12, I1, I2, k_init, corrected):
    return nrsfm.NormalsCase(pair_ptr=np.array([0, 1], np.int32), J12=J12, J21=J21, H12=H12, I1=I1, I2=I2,
                             pair_from_ref=np.ones(1, np.uint8), k_first=np.full((1, 2), np.nan, np.float32),
                             k_init=np.array([k_init], np.float64), ref_uv=I1, corrected_t2=corrected)


@pytest.mark.parametrize("seed", range(4))
def test_isometric_pair_pins_polynomials_and_transfer(orc, emu, seed):
    J12, J21, H12, I1, I2, k1, k2 = _plane_pair(seed)
    for api in (orc, emu):
        # with the transfer step's definition of t2 the true normal is a root of both polynomials ...
        nc = _normals_case(J12, J21, H12, I1, I2, k1 + 0.01, corrected=1)
        out = api.normals(nc)
        assert out.status[0] == 1
        assert np.abs(out.k[0] - k1).max() < 2e-5          # fp32 inputs
        # ... and the reference's transfer formula (NormalEstimator.cc:199-223) carries it to view 2
        assert np.abs(out.pair_normal[0, :2] - k2).max() < 2e-5
        assert abs(out.pair_normal[0, 2] - (1 - k2 @ I2[0])) < 2e-5
        # the reference's polynomial build (quirk C7: t2 ~ 0 for any projective warp) does not have
        # the true normal as a root: starting AT the truth it walks away
        nc0 = _normals_case(J12, J21, H12, I1, I2, k1, corrected=0)
        out0 = api.normals(nc0)
        assert np.abs(out0.k[0] - k1).max() > 1e-3


# --------------------------------------------------------------------------- Schwarp -------
def test_schwarzian_jacobian_matches_finite_differences(orc, win):
    c = nrsfm.schwarp_cases(win)[0]
    x0 = orc.schwarp_init(c)
    r, J = orc.schwarp_evaluate(c, x0)
    n = c.n
    rng = np.random.default_rng(0)
    for _ in range(3):
        d = rng.normal(size=len(x0)) * 1e-6
        rp, _ = orc.schwarp_evaluate(c, x0 + d, jac=False)
        rm, _ = orc.schwarp_evaluate(c, x0 - d, jac=False)
        fd, an = (rp - rm)[2 * n:] / 2, (J @ d)[2 * n:]
        assert np.abs(fd - an).max() <= 1e-7 * np.abs(an).max()


def test_data_jacobian_is_the_reference_transcription(orc, oracle, win):
    """Schwarp.cc:76-83 + :291-299: rows i and i+n are both -fx*C_i on the x columns (no 1/sigma)."""
    c = nrsfm.schwarp_cases(win)[0]
    x0 = orc.schwarp_init(c)
    _, J = orc.schwarp_evaluate(c, x0)
    n, NC = c.n, c.NC
    rc, Cm = oracle.bbs_oracle().coloc(c.bbs, c.kp1[:, 0].astype(float), c.kp1[:, 1].astype(float))
    assert rc == 0
    assert np.array_equal(J[:n, :NC], -Cm * c.fx)
    assert np.array_equal(J[n:2 * n, :NC], -Cm * c.fx)
    assert not J[:2 * n, NC:].any()
    # the residuals, on the other hand, are what the header says
    r, _ = orc.schwarp_evaluate(c, x0, jac=False)
    W = np.stack([Cm @ x0[:NC], Cm @ x0[NC:]], 1)
    assert np.allclose(r[:n], c.inv_sigma * (c.kp2[:, 0] - W[:, 0]) * c.fx, rtol=1e-12, atol=1e-12)
    assert np.allclose(r[n:2 * n], c.inv_sigma * (c.kp2[:, 1] - W[:, 1]) * c.fy, rtol=1e-12, atol=1e-12)


def test_warp_initialize_satisfies_its_normal_equations(orc, oracle, win):
    c = nrsfm.schwarp_cases(win)[0]
    x0 = orc.schwarp_init(c)
    NC = c.NC
    bo = oracle.bbs_oracle()
    _, Cm = bo.coloc(c.bbs, c.kp1[:, 0].astype(float), c.kp1[:, 1].astype(float))
    _, B = bo.bending(c.bbs)
    A = Cm.T @ Cm + c.lam * B
    for d in range(2):
        res = A @ x0[d * NC:(d + 1) * NC] - Cm.T @ c.kp2[:, d].astype(float)
        assert np.abs(res).max() < 1e-10


def test_schwarp_fit_never_increases_cost_and_counts_steps(orc, win):
    from tests import nrsfm_checks as ck
    c = nrsfm.schwarp_cases(win)[0]
    f = orc.schwarp_fit(c)
    assert f.d.iterations == 3 and f.d.cost_final <= f.d.cost_initial
    f2 = orc.schwarp_fit(ck.accepted_steps_case(c))
    assert f2.d.accepted >= 1 and f2.d.cost_final < f2.d.cost_initial


# --------------------------------------------------------------------------- SfN -----------
def test_sfn_solution_is_the_least_squares_minimiser(orc, win):
    fits = [orc.schwarp_fit(c) for c in nrsfm.schwarp_cases(win)]
    no = orc.normals(nrsfm.normals_case(win, fits))
    sc = nrsfm.sfn_case(win, no)
    A, b = orc.sfn_system(sc)
    ctrl, xyz = orc.sfn_solve(sc)
    xs = np.linalg.lstsq(A, b, rcond=None)[0]
    corr = np.float32(1) / np.sort(xs.astype(np.float32))[len(xs) // 2]
    assert np.abs(xs * float(corr) - ctrl).max() < 1e-9
    assert abs(np.sort(ctrl.astype(np.float32))[len(ctrl) // 2] - 1) < 1e-6
    # (u d, v d, d)
    assert np.allclose(xyz[:, 0], sc.eval_uv[:, 0] * xyz[:, 2], rtol=1e-6, atol=1e-7)


def test_sfn_recovers_a_plane(orc):
    """exact normals of a plane -> depths proportional to the plane's depths"""
    rng = np.random.default_rng(3)
    nrm = np.array([0.15, -0.2, 1.0])
    q = rng.uniform(-0.5, 0.5, (400, 2)).astype(np.float32)
    umin, umax, vmin, vmax = nrsfm.keyframe_domain(q)
    depth = 1.0 / (q.astype(float) @ nrm[:2] + nrm[2])
    sc = nrsfm.SfnCase(bbs=nrsfm.make_bbs(umin, umax, vmin, vmax, valdim=1), uv=q,
                       normals=np.tile(nrm.astype(np.float32), (400, 1)), eval_uv=q, bending=1e-3)
    ctrl, xyz = orc.sfn_solve(sc)
    ratio = xyz[:, 2] / depth
    assert ratio.std() / ratio.mean() < 2e-3


# --------------------------------------------------------------------------- regression ----
def test_oracle_reproduces_committed_vectors(orc, win):
    g = helpers.golden("nrsfm_oracle.npz")
    fits = []
    for i, c in enumerate(nrsfm.schwarp_cases(win)):
        f = orc.schwarp_fit(c)
        fits.append(f)
        assert np.abs(f.x - g[f"fit{i}_x"]).max() < 1e-10
        assert np.abs(f.J12 - g[f"fit{i}_J12"]).max() < 1e-6 and np.abs(f.H12 - g[f"fit{i}_H12"]).max() < 1e-4
        assert (f.keep == g[f"fit{i}_keep"]).all()
        sc = g[f"fit{i}_scalars"]
        assert abs(f.d.cost_final - sc[1]) < 1e-9 * sc[1] and f.d.iterations == sc[2] and f.d.accepted == sc[3]
    no = orc.normals(nrsfm.normals_case(win, fits))
    assert (no.status == g["normals_status"]).all()
    ok = (no.status == 1) & (np.abs(g["normals_cov"]).max(1) < 1e6) & (g["normals_iters"] < 200)
    assert np.abs(no.k - g["normals_k"])[ok[:len(no.k)]].max() < 1e-8
    ctrl, xyz = orc.sfn_solve(nrsfm.sfn_case(win, no))
    assert np.abs(ctrl - g["sfn_ctrl"]).max() < 1e-7


# --------------------------------------------------------------------------- Sim(3) --------
def test_sim3_numeric_jacobian_is_the_analytic_one(orc, oracle):
    """g2o differentiates EdgeSim3Simple numerically (delta 1e-9); the kernel uses the limit
    [y]x | -I | -y with y = S.map(p1)."""
    c = nrsfm.sim3_case(1, n=20)
    q = np.array([0.02, -0.01, 0.03, 1.0]); q /= np.linalg.norm(q)
    c.rot, c.trans, c.scale = tuple(q), (0.01, -0.02, 0.03), 1.1
    lib = oracle.load()
    lib.oracle_sim3_jacobian.restype = C.c_int
    lib.oracle_sim3_jacobian.argtypes = [C.POINTER(_capi.Sim3Problem), C.c_int, _capi.c_double_p]
    p = c.problem()
    x, y, z, w = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    for i in range(5):
        J = np.zeros((3, 7))
        lib.oracle_sim3_jacobian(C.byref(p), i, _capi.as_ptr(J, C.c_double))
        yv = c.scale * R @ c.pts1[i].astype(float) + np.array(c.trans)
        skew = np.array([[0, -yv[2], yv[1]], [yv[2], 0, -yv[0]], [-yv[1], yv[0], 0]])
        Ja = np.concatenate([skew, -np.eye(3), -yv[:, None]], 1)
        assert np.abs(J - Ja).max() < 5e-7


def test_sim3_recovers_a_noise_free_similarity(orc):
    rng = np.random.default_rng(4)
    P = rng.uniform(-0.5, 0.5, (100, 3)) + [0, 0, 1]
    R = nrsfm._rot([0.3, -0.2, 0.9], 0.04)
    s, t = 1.23, np.array([0.02, -0.03, 0.05])
    # values exactly representable in fp32 on both sides keep the residual at rounding level
    P32 = P.astype(np.float32)
    Q32 = (s * (R @ P32.astype(float).T).T + t).astype(np.float32)
    c = nrsfm.Sim3Case(pts1=P32, pts2=Q32, scale=1.0)
    r = orc.sim3_register([c])[0]
    assert abs(r["scale"] - s) < 1e-6 and np.abs(r["trans"] - t).max() < 1e-6
    assert r["acceptable"] == 1 and r["inliers"] == 100


def test_min_median_scale_is_robust_and_reproducible(orc):
    rng = np.random.default_rng(8)
    mono = rng.uniform(0.5, 1.5, (400, 3)).astype(np.float32)
    stereo = (1.7 * mono + rng.normal(size=mono.shape) * 0.002).astype(np.float32)
    stereo[:40] += 1.0                                  # gross outliers
    s1 = orc.scale_min_median(mono, stereo, seed=3)
    assert abs(s1 - 1.7) < 5e-3
    assert s1 == orc.scale_min_median(mono, stereo, seed=3)
    assert abs(orc.scale_min_median(mono, stereo, seed=4) - 1.7) < 5e-3
