"""Writes tests/golden/sft_ref.npz: outputs of the REFERENCE'S OWN SfT code (oracle/_ref/libg2o_sft_ref.so =
Thirdparty/g2o/g2o/types/sft_types.h, se3quat.h, core/base_*_edge.hpp compiled where they lie + the Levenberg
driver / Huber kernel / chi2 / update bodies extracted by line range, see oracle/g2o_ref_harness.cc) on seeded
synthetic frames.  Run in the build container (needs /root/reference):  python tests/golden/make_golden_sft.py
The vectors pin oracle/sft_oracle.c (tests/test_oracle_sft_ref.py) where the reference tree is absent."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from defslam_b200 import synthetic  # noqa: E402
from oracle import oracle_py  # noqa: E402


def cases():
    """name -> frame.  'huber' has few frame keypoints so that Omega = invSigma2/N is large and the gross
    outliers sit on the linear branch of the Huber kernel (with N = 1200 nothing ever does, quirk C4)."""
    out = {}
    for cfg in ("C1", "C2", "C4", "C3"):
        _, frames = synthetic.make_config_frames(cfg, nframes=2 if cfg == "C1" else 1)
        for i, f in enumerate(frames):
            out[f"{cfg}_{i}"] = f
    t6 = synthetic.make_template(6)
    out["tiny"] = synthetic.make_frame(t6, 40, seed=7)
    out["huber"] = synthetic.make_frame(t6, 60, seed=11, n_frame_keypoints=6, outlier_frac=0.2)
    t9 = synthetic.make_template(9)
    out["huber9"] = synthetic.make_frame(t9, 300, seed=12, n_frame_keypoints=30, outlier_frac=0.1)
    # the matches-given overload (DefOptimizer.cc:582-837): every node free, Omega = I/#matches, Huber 0.5,
    # no temporal term, curvature on the viewed nodes with the caller's lenghtEdge_ (quirk C8)
    for name, tm, M, seed in (("mg_tiny", t6, 40, 21), ("mg_9", t9, 300, 22)):
        f = synthetic.make_frame(tm, M, seed=seed, noise_px=0.2)
        f.matches_given, f.curv_edge_len = 1, float(tm.desc().edge_median_len)
        out[name] = f
    return out


def sim3_cases():
    from defslam_b200 import nrsfm
    return [nrsfm.sim3_case(s) for s in range(4)] + [nrsfm.sim3_case(9, n=40, noise=1e-4, outlier_frac=0.0)]


def mesh_cases():
    """regular grids (perturbed so that no angle is degenerate) and an irregular Delaunay mesh"""
    from scipy.spatial import Delaunay
    out = {}
    for G in (6, 9, 13):
        t = synthetic.make_template(G)
        rng = np.random.default_rng(G)
        out[f"grid{G}"] = (t.nodes_rest + 0.01 * rng.normal(size=t.nodes_rest.shape), t.facets)
    rng = np.random.default_rng(77)
    uv = rng.uniform(-0.5, 0.5, (120, 2))
    tri = Delaunay(uv)
    xyz = np.column_stack([uv, 1.0 + 0.1 * np.sin(4 * uv[:, 0]) * np.cos(3 * uv[:, 1])])
    out["delaunay"] = (xyz, tri.simplices.astype(np.int32))
    return out


def main():
    ref = oracle_py.load_g2o_ref()
    assert ref is not None, "build oracle/_ref first (make -C oracle ref)"
    g = {}
    for name, f in cases().items():
        small = name in ("tiny", "huber", "mg_tiny")
        res, J = oracle_py.sft_residuals(f, ref, "ref_sft_residuals", jac=small)
        H, b, chi = oracle_py.sft_normal_equations(f, ref, "ref_sft_normal_equations")
        o = oracle_py.sft_solve(f, ref, "ref_sft_solve")
        k = o.r.lm_iterations
        g[f"{name}.res"] = res
        g[f"{name}.b"] = b
        g[f"{name}.chi2"] = np.array(chi)
        g[f"{name}.Hdiag"] = np.diag(H).copy()
        g[f"{name}.Hcam"] = H[-6:, :].copy()
        if small:
            g[f"{name}.H"] = H
            nz = np.flatnonzero(J)
            g[f"{name}.Jnz"] = nz.astype(np.int64)
            g[f"{name}.Jval"] = J.ravel()[nz]
        g[f"{name}.nodes"] = o.nodes
        g[f"{name}.T_cw"] = o.T_cw
        g[f"{name}.outlier"] = o.outlier
        g[f"{name}.trace"] = o.trace[:k].copy()
        g[f"{name}.scalars"] = np.array([o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers, o.r.rep_error,
                                         o.r.chi2_initial, o.r.chi2_final, o.r.lambda_final])
        print(name, "rows", len(res), "its", k, "trials", o.r.lm_trials, "inliers", o.r.n_inliers, "chi2", chi,
              "->", o.r.chi2_final)
    # the pose update and the Huber kernel on their own
    rng = np.random.default_rng(5)
    q = rng.normal(size=(16, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1
    t = rng.normal(size=(16, 3))
    u = rng.normal(size=(16, 6)) * np.array([0.2, 0.2, 0.2, 0.1, 0.1, 0.1])
    u[:4, :3] *= 1e-7  # small-angle branch of SE3Quat::exp (theta < 1e-5)
    qo, to = np.zeros_like(q), np.zeros_like(t)
    import ctypes as C
    from defslam_b200 import _capi
    for i in range(16):
        a = [np.ascontiguousarray(x) for x in (q[i], t[i], u[i])]
        qq, tt = np.zeros(4), np.zeros(3)
        ref.ref_se3_oplus(*[_capi.as_ptr(x, C.c_double) for x in a], _capi.as_ptr(qq, C.c_double),
                          _capi.as_ptr(tt, C.c_double))
        qo[i], to[i] = qq, tt
    g["se3.q"], g["se3.t"], g["se3.u"], g["se3.q_out"], g["se3.t_out"] = q, t, u, qo, to
    e2 = np.concatenate([np.linspace(0, 12, 49), [5.9909, 5.99099, 5.991, 5.99101, 5.9911, 100.0]])
    rho = np.zeros((len(e2), 3))
    for i, e in enumerate(e2):
        r = np.zeros(3)
        ref.ref_huber(np.float32(np.sqrt(5.991)), float(e), _capi.as_ptr(r, C.c_double))
        rho[i] = r
    g["huber.e2"], g["huber.rho"] = e2, rho
    # LaplacianMesh::ExtractMeanCurvatures, the reference's own lines (oracle/template_ref_harness.cc)
    for name, (xyz, fac) in mesh_cases().items():
        rc, r = oracle_py.ref_mesh_laplacian(ref, xyz, fac, max_ring=16)
        assert rc == 0
        for k in ("cnt", "idx", "w", "boundary", "kappa0"):
            g[f"mesh.{name}.{k}"] = r[k]
    # Optimizer::OptimizeHorn on the reference's own Sim3 / EdgeSim3Simple / numeric Jacobians / LM driver
    rows = []
    for c in sim3_cases():
        p = c.problem()
        r = _capi.Sim3Result()
        ref.ref_sim3_optimize_horn(C.byref(p), C.byref(r))
        rows.append(list(r.rot[:]) + list(r.trans[:]) + [r.scale, r.chi2, r.inliers, r.acceptable, r.iterations[0]])
    g["sim3.out"] = np.array(rows)
    path = os.path.join(ROOT, "tests", "golden", "sft_ref.npz")
    np.savez_compressed(path, **g)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
