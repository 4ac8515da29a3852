"""GPU parity tests of the template-construction and B-spline kernels, through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from defslam_b200 import _capi, synthetic
from tests.helpers import embed_call, golden, mesh_laplacian_call

pytestmark = pytest.mark.gpu
ORDERS = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]


def _cuda_bbs(cuda_lib):
    from oracle import oracle_py as O
    return O.BbsApi(cuda_lib, "defslam_")


def test_bbs_eval_coloc_bit_exact_vs_reference_golden(cuda_lib):
    api = _cuda_bbs(cuda_lib)
    g = golden("bbs_ref.npz")
    for ci in range(4):
        dom = g[f"c{ci}_dom"]
        b = _capi.Bbs(dom[0], dom[1], int(dom[2]), dom[3], dom[4], int(dom[5]), int(dom[6]))
        for du, dv in ORDERS:
            rc, val = api.eval(b, g[f"c{ci}_ctrl"], g[f"c{ci}_u"], g[f"c{ci}_v"], du, dv)
            assert rc == 0
            ref = g[f"c{ci}_eval_{du}{dv}"]
            # nvcc contracts a*b+c into FMAs: allow one rounding of difference, no more
            assert np.abs(val - ref).max() <= 4e-16 * max(1.0, np.abs(ref).max()) * 16, (ci, du, dv)
            rc, Cm = api.coloc(b, g[f"c{ci}_u"], g[f"c{ci}_v"], du, dv)
            assert rc == 0
            rc_ref = g[f"c{ci}_coloc_{du}{dv}"]
            assert np.abs(Cm - rc_ref).max() <= 1e-15 * max(1.0, np.abs(rc_ref).max())
            assert np.array_equal(Cm != 0, rc_ref != 0)
        rc, B = api.bending(b)
        assert rc == 0
        assert np.abs(B - g[f"c{ci}_bending"]).max() <= 1e-14 * np.abs(g[f"c{ci}_bending"]).max()


def test_bbs_eval6_equals_six_single_evals(cuda_lib):
    api = _cuda_bbs(cuda_lib)
    b = _capi.Bbs(-0.9, 0.8, 13, -0.7, 0.6, 15, 2)
    rng = np.random.default_rng(3)
    ctrl = rng.normal(size=13 * 15 * 2)
    u, v = rng.uniform(b.umin, b.umax, 1000), rng.uniform(b.vmin, b.vmax, 1000)
    out = np.zeros((6, 1000, 2))
    rc = cuda_lib.defslam_bbs_eval6(C.byref(b), _capi.as_ptr(ctrl, C.c_double), 1000, _capi.as_ptr(u, C.c_double),
                                    _capi.as_ptr(v, C.c_double), _capi.as_ptr(out, C.c_double))
    assert rc == 0
    for k, (du, dv) in enumerate(ORDERS):
        assert np.array_equal(out[k], api.eval(b, ctrl, u, v, du, dv)[1])


def test_bbs_outside_domain_and_surface_vertices(cuda_lib, oracle):
    api = _cuda_bbs(cuda_lib)
    b = _capi.Bbs(-0.9, 0.8, 13, -0.7, 0.6, 15, 1)
    rc, Cm = api.coloc(b, np.array([0.0, 0.9]), np.array([0.0, 0.0]))
    assert rc == _capi.EBADARG and not Cm.any()
    ctrl = 1.0 + 0.1 * np.random.default_rng(2).normal(size=13 * 15)
    from oracle import oracle_py as O
    a = O.bbs_oracle().surface_vertices(b, ctrl, 10, 10)[1]
    rc, k = api.surface_vertices(b, ctrl, 10, 10)
    assert rc == 0
    assert np.abs(a - k).max() <= 2e-7 * np.abs(a).max()


def test_mesh_laplacian_kernel_vs_oracle(cuda_lib, oracle):
    lib = oracle.load()
    for G in (9, 13, 25):
        tmpl = synthetic.make_template(G)
        rc_o, o = mesh_laplacian_call(lib.oracle_mesh_laplacian, tmpl.nodes_rest, tmpl.facets)
        rc_k, k = mesh_laplacian_call(cuda_lib.defslam_mesh_laplacian, tmpl.nodes_rest, tmpl.facets)
        assert rc_o == 0 and rc_k == 0
        for key in ("cnt", "idx", "boundary", "edge_ab", "n_edges"):
            assert np.array_equal(o[key], k[key]), key
        assert np.allclose(o["w"], k["w"], rtol=1e-12, atol=1e-15)
        assert np.allclose(o["kappa0"], k["kappa0"], rtol=1e-11, atol=1e-15)
        assert np.allclose(o["edge_len0"], k["edge_len0"], rtol=1e-14)
        assert abs(o["median"] - k["median"]) <= 1e-14 * o["median"]


def test_mesh_laplacian_feeds_the_solver(cuda_lib, oracle):
    """template constants from the GPU kernel -> same SfT solution as with the oracle's constants"""
    from defslam_b200 import sft
    tmpl = synthetic.make_template(9)
    rc, k = mesh_laplacian_call(cuda_lib.defslam_mesh_laplacian, tmpl.nodes_rest, tmpl.facets)
    assert rc == 0
    n = tmpl.n_nodes
    ptr = np.zeros(n + 1, np.int32)
    ptr[1:] = np.cumsum(k["cnt"])
    idx = np.concatenate([k["idx"][i, :k["cnt"][i]] for i in range(n)]).astype(np.int32)
    w = np.concatenate([k["w"][i, :k["cnt"][i]] for i in range(n)])
    t2 = synthetic.MeshTemplate(tmpl.nodes_rest, tmpl.facets, ptr, idx, w, k["boundary"], k["kappa0"],
                                np.ascontiguousarray(k["edge_ab"]), np.ascontiguousarray(k["edge_len0"]),
                                k["median"], tmpl.uv, tmpl.G)
    f = synthetic.make_frame(tmpl, 300, seed=77)
    ref = oracle.sft_solve(f)
    f.template = t2
    out = sft.solve_batched([f])[0]
    # the two sets of constants differ in the last bits, which may flip one accept/reject
    # decision at the tail of the LM run; the solution itself must agree
    assert np.abs(out.nodes - ref.nodes).max() < 1e-6
    assert abs(out.r.lm_trials - ref.r.lm_trials) <= 3


def test_embed_points_bit_exact(cuda_lib, oracle):
    lib = oracle.load()
    tmpl = synthetic.make_template(13)
    rng = np.random.default_rng(9)
    uvn = rng.uniform(-0.75, 0.6, (1200, 2))
    d = synthetic.template_surface_depth(uvn[:, 0], uvn[:, 1]) + 0.002 * rng.normal(size=1200)
    pts = np.stack([uvn[:, 0] * d, uvn[:, 1] * d, d], 1).astype(np.float32)
    rc, of, on, ob = embed_call(lib.oracle_embed_points, tmpl.nodes_rest, tmpl.facets, pts)
    rc2, kf, kn, kb = embed_call(cuda_lib.defslam_embed_points, tmpl.nodes_rest, tmpl.facets, pts)
    assert rc == 0 and rc2 == 0
    assert np.array_equal(of, kf) and np.array_equal(on, kn)
    assert np.array_equal(ob, kb)          # fp32 arithmetic written with explicit roundings


def test_mappoints_recalculate(cuda_lib, oracle):
    lib = oracle.load()
    tmpl, frames = synthetic.make_config_frames("C2", nframes=1)
    f = frames[0]
    a = np.zeros((f.n_matches, 3), np.float32)
    b = np.zeros((f.n_matches, 3), np.float32)
    args = lambda out: (tmpl.n_nodes, _capi.as_ptr(f.node_xyz, C.c_double), f.n_matches,
                        _capi.as_ptr(f.match_nodes, C.c_int32), _capi.as_ptr(f.match_bary, C.c_double),
                        _capi.as_ptr(out, C.c_float))
    assert lib.oracle_mappoints_recalculate(*args(a)) == 0
    assert cuda_lib.defslam_mappoints_recalculate(*args(b)) == 0
    assert np.abs(a - b).max() <= 1.2e-7 * np.abs(a).max()
