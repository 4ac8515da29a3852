"""Pins of the CPU oracle of the SfT solve (the reference ships no tests for this path):
finite-difference Jacobians, fixed points, golden regression vectors."""
import ctypes as C

import numpy as np
import pytest

from defslam_b200 import _capi, synthetic
from tests.helpers import golden


def _residuals_and_jac(oracle, frame):
    lib = oracle.load()
    n = frame.template.n_nodes
    D = 3 * n + 6
    p = frame.problem()
    rows = lib.oracle_sft_residuals(C.byref(p), None, None, 0)
    res = np.zeros(rows)
    J = np.zeros((rows, D))
    assert lib.oracle_sft_residuals(C.byref(p), _capi.as_ptr(res, C.c_double), _capi.as_ptr(J, C.c_double), rows) == rows
    return res, J


def _residuals_at(oracle, frame, d):
    """residuals after applying the update d exactly as the solver does (nodes += d, pose = exp(dc) pose)"""
    lib = oracle.load()
    n = frame.template.n_nodes
    p = frame.problem()
    nodes = np.zeros((n, 3))
    q = np.zeros(4)
    t = np.zeros(3)
    d = np.ascontiguousarray(d)
    lib.oracle_sft_apply_update(C.byref(p), _capi.as_ptr(d, C.c_double), _capi.as_ptr(nodes, C.c_double), None,
                                _capi.as_ptr(q, C.c_double), _capi.as_ptr(t, C.c_double))
    rows = lib.oracle_sft_residuals(C.byref(p), None, None, 0)
    res = np.zeros(rows)
    lib.oracle_sft_residuals_pose(C.byref(p), _capi.as_ptr(q, C.c_double), _capi.as_ptr(t, C.c_double),
                                  _capi.as_ptr(nodes, C.c_double), _capi.as_ptr(res, C.c_double), rows)
    return res


@pytest.fixture(scope="module")
def small_frame():
    tmpl = synthetic.make_template(6)
    f = synthetic.make_frame(tmpl, 40, seed=7)
    # perturb the state so that no residual is at a degenerate point
    rng = np.random.default_rng(3)
    f.node_xyz = np.ascontiguousarray(f.node_xyz + 0.01 * rng.normal(size=f.node_xyz.shape))
    return f


def test_exact_jacobians_match_central_differences(oracle, small_frame):
    f = small_frame
    n = f.template.n_nodes
    D = 3 * n + 6
    res, J = _residuals_and_jac(oracle, f)
    nrep = 2 * f.n_matches
    h = 1e-6
    Jfd = np.zeros_like(J)
    for k in range(D):
        d = np.zeros(D)
        d[k] = h
        rp = _residuals_at(oracle, f, d)
        d[k] = -h
        rm = _residuals_at(oracle, f, d)
        Jfd[:, k] = (rp - rm) / (2 * h)
    scale = np.abs(Jfd).max()
    # temporal, curvature, stretch rows: analytic == FD for every column
    assert np.abs(J[nrep:] - Jfd[nrep:]).max() < 1e-6 * max(1.0, np.abs(Jfd[nrep:]).max())
    # reprojection rows: the camera block (last 6 columns) is exact
    assert np.abs(J[:nrep, 3 * n:] - Jfd[:nrep, 3 * n:]).max() < 1e-5 * scale
    # ... but the node block is the reference's deliberate approximation (quirk C1):
    # it linearises at each node's own camera-frame position, so it does NOT match FD
    dev = np.abs(J[:nrep, :3 * n] - Jfd[:nrep, :3 * n]).max() / np.abs(Jfd[:nrep, :3 * n]).max()
    assert 1e-3 < dev < 0.5


def test_node_jacobian_is_literal_transcription_of_the_reference_formula(oracle, small_frame):
    """J_node_k = -(1/z_k) [[fx,0,-x_k/z_k fx],[0,fy,-y_k/z_k fy]] R b_k with (x_k,y_k,z_k) = R x_k + t
    (sft_types.h:176-205)."""
    f = small_frame
    n = f.template.n_nodes
    res, J = _residuals_and_jac(oracle, f)
    T = f.T_cw.astype(np.float64)
    R, t = T[:3, :3], T[:3, 3]
    for m in range(0, f.n_matches, 7):
        for k in range(3):
            v = f.match_nodes[m, k]
            xc = R @ f.node_xyz[v] + t
            tmp = np.array([[f.fx, 0, -xc[0] / xc[2] * f.fx], [0, f.fy, -xc[1] / xc[2] * f.fy]])
            Jn = -1.0 / xc[2] * tmp @ R * f.match_bary[m, k]
            assert np.allclose(J[2 * m:2 * m + 2, 3 * v:3 * v + 3], Jn, rtol=1e-9, atol=1e-12)


def test_normal_equations_are_jt_w_j(oracle, small_frame):
    """H and b of the oracle == J^T W J, -J^T W r assembled from its own residual Jacobian."""
    f = small_frame
    n = f.template.n_nodes
    res, J = _residuals_and_jac(oracle, f)
    H, b, chi = oracle.sft_normal_equations(f)
    out = oracle.sft_solve(f)
    role = out.role
    viewed, optlap = (role & 1).astype(bool), (role >> 1).astype(bool)
    N = f.n_frame_keypoints
    nrep = 2 * f.n_matches
    w = np.zeros(len(res))
    info = f.match_inv_sigma2.astype(np.float64) / N
    delta = float(np.float32(np.sqrt(5.991)))
    for m in range(f.n_matches):
        c2 = info[m] * (res[2 * m] ** 2 + res[2 * m + 1] ** 2)
        rho1 = 1.0 if c2 <= delta * delta else delta / np.sqrt(c2)
        w[2 * m:2 * m + 2] = info[m] * rho1
    nv = int(viewed.sum())
    w[nrep:nrep + 3 * nv] = f.reg_temp / f.template.edge_median_len ** 2
    ncurv = len(res) - nrep - 3 * nv
    # remaining rows: curvature then stretch; their counts follow from the graph
    interior = optlap & (f.template.boundary == 0)
    deg = np.diff(np.r_[0, np.cumsum(np.bincount(f.template.edge_ab.reshape(-1), minlength=n))])
    n_curv = int(deg[interior].sum())
    n_str = ncurv - n_curv
    w[nrep + 3 * nv:nrep + 3 * nv + n_curv] = f.reg_lap / optlap.sum()
    w[nrep + 3 * nv + n_curv:] = f.reg_inex / n_str
    free = np.r_[np.repeat(optlap, 3), np.ones(6, bool)]
    Jf = J[:, free]
    Href = Jf.T @ (w[:, None] * Jf)
    bref = -Jf.T @ (w * res)
    assert np.allclose(H[np.ix_(free, free)], Href, rtol=1e-9, atol=1e-12 * np.abs(Href).max())
    assert np.allclose(b[free], bref, rtol=1e-9, atol=1e-12 * np.abs(bref).max())


def test_rest_shape_with_exact_observations_is_a_fixed_point(oracle):
    tmpl = synthetic.make_template(7)
    f = synthetic.make_frame(tmpl, 120, seed=11, noise_px=0.0, outlier_frac=0.0, amp=0.0, shear=0.0, rot_deg=0.0,
                             trans=0.0)
    out = oracle.sft_solve(f)
    # float32 observations leave a tiny reprojection residual; nothing else pulls on the nodes
    assert out.r.chi2_initial < 1e-6
    assert np.abs(out.nodes - tmpl.nodes_rest).max() < 1e-5
    assert np.abs(out.T_cw - np.eye(4)).max() < 1e-5
    assert out.r.n_inliers == f.n_matches


def test_solution_reduces_cost_and_flags_are_consistent(oracle):
    tmpl, frames = synthetic.make_config_frames("C1", nframes=2)
    for f in frames:
        o = oracle.sft_solve(f)
        assert o.r.chi2_final < o.r.chi2_initial
        assert 1 <= o.r.lm_iterations <= 50 and o.r.lm_trials >= o.r.lm_iterations
        tr = o.trace[:o.r.lm_iterations]
        assert np.all(tr[:, 3] <= tr[:, 0] + 1e-12)          # accepted steps never increase chi2
        assert np.all(np.diff(tr[:, 0]) <= 1e-12)
        viewed = (o.role & 1).astype(bool)
        assert set(np.flatnonzero(viewed)) == set(np.unique(f.match_nodes))
        assert np.all(((o.role >> 1) & 1)[viewed] == 1)


def test_oracle_matches_committed_golden_vectors(oracle):
    g = golden("sft_oracle.npz")
    for cfg, nfr in [("C1", 3), ("C4", 2), ("C2", 1)]:
        tmpl, frames = synthetic.make_config_frames(cfg, nframes=nfr)
        for i, f in enumerate(frames):
            o = oracle.sft_solve(f)
            k = f"{cfg}_{i}"
            assert np.allclose(o.nodes, g[k + "_nodes"], rtol=0, atol=1e-10)
            assert np.array_equal(o.outlier[:f.n_matches], g[k + "_outlier"])
            assert o.r.lm_iterations == int(g[k + "_scalars"][0]) and o.r.lm_trials == int(g[k + "_scalars"][1])
            H, b, chi = oracle.sft_normal_equations(f)
            assert np.allclose(np.diag(H), g[k + "_Hdiag"], rtol=1e-12)
            assert np.allclose(b, g[k + "_b"], rtol=1e-10, atol=1e-14)
