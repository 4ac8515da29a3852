"""The SfT oracle pinned to the REFERENCE'S OWN CODE.

oracle/_ref/libg2o_sft_ref.so is DefSLAM's own sft_types.h / se3quat.h / base_{unary,binary,multi}_edge.hpp compiled
where they lie, plus the Levenberg-Marquardt driver (optimization_algorithm_levenberg.cpp:43-189), the Huber kernel
(robust_kernel_impl.cpp:65-91), activeRobustChi2 / update (sparse_optimizer.cpp:104-120,477-491) and the vertex oplus
bodies extracted by line range (oracle/g2o_ref_harness.cc, oracle/ref_shim/).  Two tiers:
  * golden: tests/golden/sft_ref.npz holds that library's outputs (tests/golden/make_golden_sft.py) -- runs anywhere;
  * live:   the same comparison against the library itself on C1-C5 frames -- runs where oracle/_ref/ was built.
Tolerances: per-edge errors and Jacobians 1e-15 (they are the same expressions), H/b/chi2 1e-13 (summation order),
LM trace 1e-9, final nodes 1e-9 relative; iteration / trial counts and outlier flags identical."""
import ctypes as C

import numpy as np
import pytest

from defslam_b200 import _capi, synthetic
from tests.helpers import golden
from tests.golden.make_golden_sft import cases


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


@pytest.fixture(scope="module")
def gold():
    return golden("sft_ref.npz")


@pytest.fixture(scope="module")
def frames():
    return cases()


@pytest.fixture(scope="module")
def ref(oracle):
    lib = oracle.load_g2o_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libg2o_sft_ref.so not built here (needs the reference tree)")
    return lib


NAMES = ["C1_0", "C1_1", "C2_0", "C4_0", "C3_0", "tiny", "huber", "huber9", "mg_tiny", "mg_9"]


@pytest.mark.parametrize("name", NAMES)
def test_edge_errors_and_normal_equations_equal_reference_golden(oracle, gold, frames, name):
    f = frames[name]
    small = name in ("tiny", "huber", "mg_tiny")
    res, J = oracle.sft_residuals(f, jac=small)
    assert res.shape == gold[f"{name}.res"].shape
    assert _rel(res, gold[f"{name}.res"]) < 1e-15
    if small:
        Jg = np.zeros(J.size)
        Jg[gold[f"{name}.Jnz"]] = gold[f"{name}.Jval"]
        assert _rel(J.ravel(), Jg) < 1e-15
    H, b, chi = oracle.sft_normal_equations(f)
    assert abs(chi - float(gold[f"{name}.chi2"])) <= 1e-13 * abs(chi)
    assert _rel(b, gold[f"{name}.b"]) < 1e-13
    assert _rel(np.diag(H), gold[f"{name}.Hdiag"]) < 1e-13
    assert _rel(H[-6:], gold[f"{name}.Hcam"]) < 1e-13
    if small:
        assert _rel(H, gold[f"{name}.H"]) < 1e-13


@pytest.mark.parametrize("name", NAMES)
def test_lm_solve_equals_reference_golden(oracle, gold, frames, name):
    f = frames[name]
    o = oracle.sft_solve(f)
    its, trials, inl, rep, chi0, chi1, lam = gold[f"{name}.scalars"]
    assert (o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers) == (int(its), int(trials), int(inl))
    k = o.r.lm_iterations
    tr = gold[f"{name}.trace"]
    assert np.array_equal(o.trace[:k, 2], tr[:, 2])                 # trials per iteration
    assert _rel(o.trace[:k, 0], tr[:, 0]) < 1e-9                     # chi2 at the start of each iteration
    assert _rel(o.trace[:k, 1], tr[:, 1]) < 1e-9                     # lambda at the start of each iteration
    assert _rel(o.trace[:k, 3], tr[:, 3]) < 1e-9                     # chi2 after each iteration
    assert abs(o.r.chi2_final - chi1) < 1e-9 * chi1 and abs(o.r.lambda_final - lam) < 1e-9 * lam
    scale = np.sqrt((gold[f"{name}.nodes"] ** 2).sum(1).mean())
    assert np.abs(o.nodes - gold[f"{name}.nodes"]).max() / scale < 1e-9
    assert np.array_equal(o.outlier[:f.n_matches], gold[f"{name}.outlier"][:f.n_matches])
    assert np.abs(o.T_cw - gold[f"{name}.T_cw"]).max() < 1e-6
    assert abs(o.r.rep_error - rep) < 1e-5 * rep


def test_huber_cases_exercise_the_linear_branch(gold):
    # with N = 1200 frame keypoints no residual ever leaves the quadratic zone (quirk C4); the two huber cases do
    for name in ("huber", "huber9"):
        assert gold[f"{name}.outlier"].sum() > 0


def test_pose_update_equals_reference_se3quat(oracle, gold):
    """oracle pose update == SE3Quat::exp(update) * estimate (se3quat.h:223-257, types_six_dof_expmap.h:73-76),
    including the small-angle branch; checked through oracle_sft_apply_update on a one-node problem."""
    lib = oracle.load()
    tmpl = synthetic.make_template(4)
    f = synthetic.make_frame(tmpl, 8, seed=1)
    n = tmpl.n_nodes
    q, t, u = gold["se3.q"], gold["se3.t"], gold["se3.u"]
    for i in range(q.shape[0]):
        # T_cw is f32 in the ABI: round the pose to what the f32 matrix holds, then compare in that setting
        x, y, z, w = q[i]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        T = np.eye(4, dtype=np.float32)
        T[:3, :3], T[:3, 3] = R, t[i]
        f.T_cw = T
        p = f.problem()
        d = np.zeros(3 * n + 6)
        d[3 * n:] = u[i]
        nodes = np.zeros((n, 3))
        qo, to = np.zeros(4), np.zeros(3)
        lib.oracle_sft_apply_update(C.byref(p), _capi.as_ptr(d, C.c_double), _capi.as_ptr(nodes, C.c_double), None,
                                    _capi.as_ptr(qo, C.c_double), _capi.as_ptr(to, C.c_double))
        # the golden output started from the unrounded pose: tolerance = f32 rounding of the input pose
        assert np.abs(qo - gold["se3.q_out"][i]).max() < 5e-7
        assert np.abs(to - gold["se3.t_out"][i]).max() < 5e-6


def test_huber_kernel_equals_reference_golden(oracle, gold):
    """rho, rho' of the oracle (through chi2 of a one-edge problem is overkill): restate delta/dsqr as the oracle
    sets them and compare with the reference's robustify, incl. the float dsqr threshold (robust_kernel_impl.h:84)."""
    delta = float(np.float32(np.sqrt(5.991)))
    dsqr = float(np.float32(delta * delta))
    for e2, rho in zip(gold["huber.e2"], gold["huber.rho"]):
        if e2 <= dsqr:
            r0, r1 = e2, 1.0
        else:
            r0, r1 = 2 * np.sqrt(e2) * delta - dsqr, delta / np.sqrt(e2)
        assert abs(r0 - rho[0]) <= 1e-15 * max(1.0, abs(rho[0])) and abs(r1 - rho[1]) <= 1e-15


def _check_mesh(o, r):
    assert np.array_equal(o["cnt"], r["cnt"]) and np.array_equal(o["boundary"], r["boundary"])
    for i in range(len(o["cnt"])):
        a = dict(zip(o["idx"][i][:o["cnt"][i]], o["w"][i]))
        b = dict(zip(r["idx"][i][:r["cnt"][i]], r["w"][i]))
        assert a.keys() == b.keys()
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-15 * abs(b[k])
    assert np.abs(o["kappa0"] - r["kappa0"]).max() <= 1e-15 * np.abs(r["kappa0"]).max()


@pytest.mark.parametrize("name", ["grid6", "grid9", "grid13", "delaunay"])
def test_mesh_laplacian_equals_reference_golden(oracle, gold, name):
    """oracle/template_oracle.c (mean-value weights, boundary flags, kappa0) against the reference's own
    LaplacianMesh::ExtractMeanCurvatures lines (LaplacianMesh.cc:53-148,157-162)"""
    from tests.golden.make_golden_sft import mesh_cases
    from tests.helpers import mesh_laplacian_call
    xyz, fac = mesh_cases()[name]
    rc, o = mesh_laplacian_call(oracle.load().oracle_mesh_laplacian, xyz, fac, max_ring=16)
    assert rc == 0
    _check_mesh(o, {k: gold[f"mesh.{name}.{k}"] for k in ("cnt", "idx", "w", "boundary", "kappa0")})


def test_live_mesh_laplacian(oracle, ref):
    from tests.helpers import mesh_laplacian_call
    for G in (10, 17, 25):
        t = synthetic.make_template(G)
        rng = np.random.default_rng(100 + G)
        xyz = t.nodes_rest + 0.01 * rng.normal(size=t.nodes_rest.shape)
        rc, o = mesh_laplacian_call(oracle.load().oracle_mesh_laplacian, xyz, t.facets)
        rc2, r = oracle.ref_mesh_laplacian(ref, xyz, t.facets)
        assert rc == 0 and rc2 == 0 and r["n_bad"] == 0
        _check_mesh(o, r)


def _check_sim3(o, row):
    # numeric Jacobians (delta 1e-9) put ~1e-10 of noise on the estimate; the SECOND run starts at the optimum,
    # where that noise decides how many iterations it takes -- not compared
    assert o["iterations"][0] == int(row[11])
    assert np.abs(o["rot"] - row[0:4]).max() < 1e-8 and np.abs(o["trans"] - row[4:7]).max() < 1e-8
    assert abs(o["scale"] - row[7]) < 1e-8 * row[7]
    assert abs(o["chi2"] - row[8]) < 1e-8 * row[8]
    assert (o["inliers"], o["acceptable"]) == (int(row[9]), int(row[10]))


def test_sim3_registration_equals_reference_golden(oracle, gold):
    """oracle/sim3_oracle.c against Optimizer::OptimizeHorn run on the reference's own sim3.h, EdgeSim3Simple,
    VertexSim3ExpmapNoProj (types_seven_dof_expmap.h:96-126,159-188), base_unary_edge.hpp numeric Jacobians,
    Huber kernel and Levenberg driver"""
    from defslam_b200 import nrsfm
    from tests.golden.make_golden_sft import sim3_cases
    orc = nrsfm.Api(oracle.load(), "oracle_")
    for o, row in zip(orc.sim3_register(sim3_cases()), gold["sim3.out"]):
        _check_sim3(o, row)


def test_live_sim3_registration(oracle, ref):
    from defslam_b200 import nrsfm
    orc = nrsfm.Api(oracle.load(), "oracle_")
    cases = [nrsfm.sim3_case(s, n=300) for s in range(20, 26)]
    for c, o in zip(cases, orc.sim3_register(cases)):
        p = c.problem()
        r = _capi.Sim3Result()
        assert ref.ref_sim3_optimize_horn(C.byref(p), C.byref(r)) == 0
        _check_sim3(o, list(r.rot[:]) + list(r.trans[:]) + [r.scale, r.chi2, r.inliers, r.acceptable, r.iterations[0]])


# ------------------------------------------------------------------ live tier (reference library present)
LIVE = [("C1", 0), ("C1", 1), ("C2", 0), ("C2", 1), ("C4", 0), ("C4", 1), ("C3", 0)]


@pytest.mark.parametrize("cfg", ["C1", "C4", "C2"])
def test_live_matches_given_overload(oracle, ref, cfg):
    """DefPoseOptimization(matches, ...) (DefOptimizer.cc:582-837) on the reference's own edges and LM driver"""
    tmpl, fr = synthetic.make_config_frames(cfg, nframes=1)
    f = fr[0]
    f.matches_given, f.curv_edge_len = 1, float(tmpl.desc().edge_median_len)
    res, J = oracle.sft_residuals(f)
    res_r, J_r = oracle.sft_residuals(f, ref, "ref_sft_residuals")
    assert _rel(res, res_r) < 1e-15 and _rel(J, J_r) < 1e-15
    H, b, chi = oracle.sft_normal_equations(f)
    Hr, br, chir = oracle.sft_normal_equations(f, ref, "ref_sft_normal_equations")
    assert _rel(H, Hr) < 1e-13 and _rel(b, br) < 1e-13 and abs(chi - chir) < 1e-13 * chir
    o = oracle.sft_solve(f)
    r = oracle.sft_solve(f, ref, "ref_sft_solve")
    assert (o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers) == (r.r.lm_iterations, r.r.lm_trials, r.r.n_inliers)
    scale = np.sqrt((r.nodes ** 2).sum(1).mean())
    assert np.abs(o.nodes - r.nodes).max() / scale < 1e-9
    assert np.array_equal(o.outlier, r.outlier)


@pytest.mark.parametrize("cfg,idx", LIVE + [("C5", 0)])
def test_live_edges_and_normal_equations(oracle, ref, cfg, idx):
    _, fr = synthetic.make_config_frames(cfg, nframes=idx + 1)
    f = fr[idx]
    jac = cfg != "C5"
    res, J = oracle.sft_residuals(f, jac=jac)
    res_r, J_r = oracle.sft_residuals(f, ref, "ref_sft_residuals", jac=jac)
    assert _rel(res, res_r) < 1e-15
    if jac:
        assert _rel(J, J_r) < 1e-15
    H, b, chi = oracle.sft_normal_equations(f)
    Hr, br, chir = oracle.sft_normal_equations(f, ref, "ref_sft_normal_equations")
    assert _rel(H, Hr) < 1e-13 and _rel(b, br) < 1e-13 and abs(chi - chir) < 1e-13 * chir


@pytest.mark.parametrize("cfg,idx", LIVE)
def test_live_lm_solve(oracle, ref, cfg, idx):
    _, fr = synthetic.make_config_frames(cfg, nframes=idx + 1)
    f = fr[idx]
    o = oracle.sft_solve(f)
    r = oracle.sft_solve(f, ref, "ref_sft_solve")
    assert (o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers) == (r.r.lm_iterations, r.r.lm_trials, r.r.n_inliers)
    k = o.r.lm_iterations
    assert np.array_equal(o.trace[:k, 2], r.trace[:k, 2])
    for c in (0, 1, 3):
        assert _rel(o.trace[:k, c], r.trace[:k, c]) < 1e-9
    scale = np.sqrt((r.nodes ** 2).sum(1).mean())
    assert np.abs(o.nodes - r.nodes).max() / scale < 1e-9
    assert np.array_equal(o.outlier, r.outlier)
    assert np.array_equal(o.T_cw, r.T_cw)
