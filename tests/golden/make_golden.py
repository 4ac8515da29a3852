"""Generates the committed golden fixtures.

  bbs_ref.npz        outputs of the REFERENCE's own Thirdparty/BBS/bbs.cc (compiled where it lies
                     into oracle/_ref/libbbs_ref.so): eval for all six derivative orders,
                     collocation rows, bending matrix.  This is a true reference pin.
  sft_oracle.npz     outputs of the CPU oracle on seeded synthetic frames (regression pin: the
                     reference ships no vectors for the SfT path and cannot be built here).
  template_oracle.npz mesh Laplacian constants of the oracle for the 9x9 synthetic template.

  polysolver_ref.npz outputs of the REFERENCE's own PolySolver::getCoefficients (PolySolver.cc:50-149,
                     compiled into oracle/_ref/libpolysolver_ref.so).  True reference pin.
  nrsfm_oracle.npz   outputs of the CPU oracle for one seeded keyframe window through the three NRSfM
                     stages (regression pin: the reference solves with Ceres/Eigen, absent here).

Run from the repo root, in the build container (needs /root/reference for the BBS library):
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from defslam_b200 import synthetic  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def bbs_case(seed, nu, nv, valdim, nsites):
    rng = np.random.default_rng(seed)
    umin, umax = -0.9 - 0.1 * rng.random(), 0.7 + 0.1 * rng.random()
    vmin, vmax = -0.7 - 0.1 * rng.random(), 0.6 + 0.1 * rng.random()
    ctrl = rng.normal(size=nu * nv * valdim)
    u = rng.uniform(umin, umax, nsites)
    v = rng.uniform(vmin, vmax, nsites)
    u[0], v[1], u[2], v[3] = umax, vmax, umin, vmin  # domain edges
    return (umin, umax, nu, vmin, vmax, nv, valdim), ctrl, u, v


def poly_inputs(n=256, seed=5):
    rng = np.random.default_rng(seed)
    J12 = (np.eye(2).reshape(1, 4) + rng.normal(size=(n, 4)) * 0.2).astype(np.float32)
    H12 = (rng.normal(size=(n, 6)) * 0.3).astype(np.float32)
    I1 = rng.uniform(-0.7, 0.7, (n, 2)).astype(np.float32)
    I2 = rng.uniform(-0.7, 0.7, (n, 2)).astype(np.float32)
    return J12, H12, I1, I2


def poly_scalars(J12, H12, I1, I2):
    """the fp32 pre-computation of NormalEstimator.cc:88-103 (t1, t2, e1, e2), numpy float32"""
    f = np.float32
    a, b, c, d = (J12[:, k] for k in range(4))
    t1 = (-b * H12[:, 4] / f(2)) + (a * H12[:, 5] / f(2))
    t2 = (-(d * H12[:, 4]) / f(2)) + ((c * H12[:, 5]) / f(2))
    e1 = (f(1) + I1[:, 0] * I1[:, 0]) + I1[:, 1] * I1[:, 1]
    e2 = (f(1) + I2[:, 0] * I2[:, 0]) + I2[:, 1] * I2[:, 1]
    return t1.astype(f), t2.astype(f), e1.astype(f), e2.astype(f)


def poly_golden():
    import ctypes as C
    from defslam_b200 import _capi
    ref = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpolysolver_ref.so"))
    ref.ref_polysolver_coefficients.argtypes = [C.c_double] * 12 + [_capi.c_double_p] * 2
    J12, H12, I1, I2 = poly_inputs()
    t1, t2, e1, e2 = poly_scalars(J12, H12, I1, I2)
    n = len(J12)
    q1, q2 = np.zeros((n, 10)), np.zeros((n, 10))
    for i in range(n):
        ref.ref_polysolver_coefficients(J12[i, 0], J12[i, 1], J12[i, 2], J12[i, 3], t1[i], t2[i], e1[i], e2[i],
                                        I1[i, 0], I1[i, 1], I2[i, 0], I2[i, 1],
                                        _capi.as_ptr(q1[i:i + 1], C.c_double), _capi.as_ptr(q2[i:i + 1], C.c_double))
    np.savez_compressed(os.path.join(OUT, "polysolver_ref.npz"), J12=J12, H12=H12, I1=I1, I2=I2, eq1=q1, eq2=q2)


def nrsfm_golden():
    from defslam_b200 import nrsfm
    api = nrsfm.Api(O.load(), "oracle_")
    win = nrsfm.make_window(42, n_keypoints=300, n_views=2)
    cases = nrsfm.schwarp_cases(win)
    d = {}
    fits = []
    for i, c in enumerate(cases):
        f = api.schwarp_fit(c)
        fits.append(f)
        d[f"fit{i}_x"] = f.x
        d[f"fit{i}_J12"], d[f"fit{i}_H12"], d[f"fit{i}_keep"] = f.J12, f.H12, f.keep
        d[f"fit{i}_scalars"] = np.array([f.d.cost_initial, f.d.cost_final, f.d.iterations, f.d.accepted])
    nc = nrsfm.normals_case(win, fits)
    no = api.normals(nc)
    d["normals_k"], d["normals_status"], d["normals_iters"] = no.k, no.status, no.iters
    d["normals_cov"], d["pair_normal"] = no.cov, no.pair_normal
    sc = nrsfm.sfn_case(win, no)
    ctrl, xyz = api.sfn_solve(sc)
    d["sfn_ctrl"], d["sfn_xyz"] = ctrl, xyz
    np.savez_compressed(os.path.join(OUT, "nrsfm_oracle.npz"), **d)


def main():
    O.build_ref()
    ref = O.BbsReference()
    data = {}
    for ci, (nu, nv, vd, ns) in enumerate([(13, 15, 2, 64), (13, 15, 1, 48), (9, 9, 2, 32), (17, 17, 1, 32)]):
        dom, ctrl, u, v = bbs_case(100 + ci, nu, nv, vd, ns)
        b = O._bbs_struct(*dom)
        data[f"c{ci}_dom"] = np.array(dom, dtype=np.float64)
        data[f"c{ci}_ctrl"], data[f"c{ci}_u"], data[f"c{ci}_v"] = ctrl, u, v
        for du, dv in [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]:
            data[f"c{ci}_eval_{du}{dv}"] = ref.eval(b, ctrl, u, v, du, dv)[1]
            data[f"c{ci}_coloc_{du}{dv}"] = ref.coloc(b, u, v, du, dv)[1].astype(np.float64)
        data[f"c{ci}_bending"] = ref.bending(b)[1]
    np.savez_compressed(os.path.join(OUT, "bbs_ref.npz"), **data)

    sft = {}
    for cfg, nfr in [("C1", 3), ("C4", 2), ("C2", 1)]:
        tmpl, frames = synthetic.make_config_frames(cfg, nframes=nfr)
        for i, f in enumerate(frames):
            o = O.sft_solve(f)
            k = f"{cfg}_{i}"
            sft[k + "_nodes"] = o.nodes
            sft[k + "_Tcw"] = o.T_cw
            sft[k + "_outlier"] = o.outlier[:f.n_matches].copy()
            sft[k + "_trace"] = o.trace[:o.r.lm_iterations].copy()
            sft[k + "_scalars"] = np.array([o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers, o.r.chi2_initial,
                                            o.r.chi2_final, o.r.rep_error, o.r.lambda_final])
            H, b, chi = O.sft_normal_equations(f)
            sft[k + "_Hdiag"] = np.diag(H).copy()
            sft[k + "_b"] = b
            sft[k + "_chi"] = np.array([chi])
    np.savez_compressed(os.path.join(OUT, "sft_oracle.npz"), **sft)

    tmpl = synthetic.make_template(9)
    np.savez_compressed(os.path.join(OUT, "template_oracle.npz"), nodes=tmpl.nodes_rest, facets=tmpl.facets,
                        nbr_ptr=tmpl.nbr_ptr, nbr_idx=tmpl.nbr_idx, nbr_w=tmpl.nbr_w, boundary=tmpl.boundary,
                        kappa0=tmpl.kappa0, edge_ab=tmpl.edge_ab, edge_len0=tmpl.edge_len0,
                        median=np.array([tmpl.edge_median_len]))
    poly_golden()
    nrsfm_golden()
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
