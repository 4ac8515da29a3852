"""Generates the committed golden fixtures.

  bbs_ref.npz        outputs of the REFERENCE's own Thirdparty/BBS/bbs.cc (compiled where it lies
                     into oracle/_ref/libbbs_ref.so): eval for all six derivative orders,
                     collocation rows, bending matrix.  This is a true reference pin.
  sft_oracle.npz     outputs of the CPU oracle on seeded synthetic frames (regression pin: the
                     reference ships no vectors for the SfT path and cannot be built here).
  template_oracle.npz mesh Laplacian constants of the oracle for the 9x9 synthetic template.

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
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
