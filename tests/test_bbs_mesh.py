"""Bicubic B-spline and template-construction device functions (emulated kernel sources) and their
oracle, against the REFERENCE's own compiled bbs.cc (oracle/_ref) and the golden vectors generated
from it."""
import ctypes as C
import os

import numpy as np
import pytest

from defslam_b200 import _capi, synthetic
from tests.helpers import embed_call, emu_lib, golden, mesh_laplacian_call

ORDERS = [(0, 0), (1, 0), (0, 1), (2, 0), (1, 1), (0, 2)]


def _apis(oracle):
    from oracle import oracle_py as O
    return {"oracle": O.bbs_oracle(), "kernel-source": O.BbsApi(emu_lib(), "emu_")}


def _cases():
    g = golden("bbs_ref.npz")
    for ci in range(4):
        dom = g[f"c{ci}_dom"]
        b = _capi.Bbs(dom[0], dom[1], int(dom[2]), dom[3], dom[4], int(dom[5]), int(dom[6]))
        yield ci, g, b


def test_eval_and_coloc_are_bit_exact_against_reference_golden(oracle):
    for name, api in _apis(oracle).items():
        for ci, g, b in _cases():
            for du, dv in ORDERS:
                rc, val = api.eval(b, g[f"c{ci}_ctrl"], g[f"c{ci}_u"], g[f"c{ci}_v"], du, dv)
                assert rc == 0
                assert np.array_equal(val, g[f"c{ci}_eval_{du}{dv}"]), (name, ci, du, dv)
                rc, Cm = api.coloc(b, g[f"c{ci}_u"], g[f"c{ci}_v"], du, dv)
                assert rc == 0
                assert np.array_equal(Cm, g[f"c{ci}_coloc_{du}{dv}"]), (name, ci, du, dv)


def test_bending_matches_reference_golden(oracle):
    for name, api in _apis(oracle).items():
        for ci, g, b in _cases():
            rc, B = api.bending(b)
            ref = g[f"c{ci}_bending"]
            assert rc == 0
            assert np.abs(B - ref).max() <= 4e-15 * np.abs(ref).max(), (name, ci)
            assert np.allclose(B, B.T, rtol=0, atol=1e-12 * np.abs(ref).max())


def test_live_reference_library_when_present(oracle):
    from oracle import oracle_py as O
    try:
        ref = O.BbsReference()
    except FileNotFoundError:
        pytest.skip("oracle/_ref/libbbs_ref.so not built")
    rng = np.random.default_rng(77)
    b = O._bbs_struct(-1.03, 0.81, 13, -0.77, 0.69, 15, 2)
    ctrl = rng.normal(size=13 * 15 * 2)
    u, v = rng.uniform(b.umin, b.umax, 300), rng.uniform(b.vmin, b.vmax, 300)
    for api in _apis(oracle).values():
        for du, dv in ORDERS:
            assert np.array_equal(api.eval(b, ctrl, u, v, du, dv)[1], ref.eval(b, ctrl, u, v, du, dv)[1])


def test_spline_identities(oracle):
    """partition of unity, derivative rows sum to zero, cubic reproduction, bending null space"""
    api = _apis(oracle)["kernel-source"]
    b = _capi.Bbs(-0.9, 0.8, 13, -0.7, 0.6, 15, 1)
    rng = np.random.default_rng(5)
    u, v = rng.uniform(b.umin, b.umax, 100), rng.uniform(b.vmin, b.vmax, 100)
    assert np.allclose(api.coloc(b, u, v, 0, 0)[1].sum(1), 1.0, atol=1e-14)
    for du, dv in ORDERS[1:]:
        assert np.abs(api.coloc(b, u, v, du, dv)[1].sum(1)).max() < 1e-9
    # control points sampled from an affine function reproduce it (Greville abscissae of a uniform cubic
    # B-spline are the knot averages: ctrl index i sits at umin + (i-1)*h)
    hu, hv = (b.umax - b.umin) / (b.nptsu - 3), (b.vmax - b.vmin) / (b.nptsv - 3)
    gu = b.umin + (np.arange(b.nptsu) - 1) * hu
    gv = b.vmin + (np.arange(b.nptsv) - 1) * hv
    ctrl = (2.0 + 0.5 * gu[:, None] - 0.25 * gv[None, :]).reshape(-1)
    val = api.eval(b, ctrl, u, v)[1][:, 0]
    assert np.allclose(val, 2.0 + 0.5 * u - 0.25 * v, atol=1e-13)
    B = api.bending(b)[1]
    assert np.abs(B @ ctrl).max() < 1e-9 * np.abs(B).max()      # affine functions have no bending energy
    assert np.linalg.eigvalsh(B).min() > -1e-9 * np.abs(B).max()


def test_sites_outside_the_domain(oracle):
    for api in _apis(oracle).values():
        b = _capi.Bbs(-0.9, 0.8, 13, -0.7, 0.6, 15, 1)
        u, v = np.array([0.0, 0.9]), np.array([0.0, 0.0])
        rc, Cm = api.coloc(b, u, v)
        assert rc == _capi.EBADARG and not Cm.any()          # the reference aborts leaving zeros
        rc, val = api.eval(b, np.ones(13 * 15), u, v)
        assert np.isfinite(val[0, 0]) and np.isnan(val[1, 0])


def test_surface_vertices(oracle):
    apis = _apis(oracle)
    b = _capi.Bbs(-0.9, 0.8, 13, -0.7, 0.6, 15, 1)
    ctrl = 1.0 + 0.1 * np.random.default_rng(2).normal(size=13 * 15)
    a = apis["oracle"].surface_vertices(b, ctrl, 10, 10)[1]
    k = apis["kernel-source"].surface_vertices(b, ctrl, 10, 10)[1]
    assert np.array_equal(a, k)
    assert np.allclose(a[:, 2], a[:, 2].clip(0.5, 1.5))


def test_mesh_laplacian_kernel_source_vs_oracle_and_golden(oracle):
    lib = oracle.load()
    g = golden("template_oracle.npz")
    rc_o, o = mesh_laplacian_call(lib.oracle_mesh_laplacian, g["nodes"], g["facets"])
    rc_e, e = mesh_laplacian_call(emu_lib().emu_mesh_laplacian, g["nodes"], g["facets"])
    assert rc_o == 0 and rc_e == 0
    for key in ("cnt", "idx", "boundary", "edge_ab", "n_edges"):
        assert np.array_equal(o[key], e[key]), key
    for key in ("w", "kappa0", "edge_len0"):
        assert np.array_equal(o[key], e[key]), key
    assert o["median"] == e["median"] == float(g["median"][0])
    # against the independent NumPy restatement that produced the golden file
    ptr = g["nbr_ptr"]
    for i in range(len(ptr) - 1):
        assert np.array_equal(e["idx"][i, :e["cnt"][i]], g["nbr_idx"][ptr[i]:ptr[i + 1]])
        assert np.allclose(e["w"][i, :e["cnt"][i]], g["nbr_w"][ptr[i]:ptr[i + 1]], rtol=1e-13)
    assert np.array_equal(e["boundary"], g["boundary"])
    assert np.allclose(e["kappa0"], g["kappa0"], rtol=1e-12, atol=1e-16)
    assert np.array_equal(e["edge_ab"], g["edge_ab"])


def test_mesh_laplacian_ring_overflow_and_bad_facets(oracle):
    g = golden("template_oracle.npz")
    rc, _ = mesh_laplacian_call(emu_lib().emu_mesh_laplacian, g["nodes"], g["facets"], max_ring=4)
    assert rc == _capi.ETOOLARGE
    bad = g["facets"].copy()
    bad[0, 0] = bad[0, 1]
    rc, _ = mesh_laplacian_call(emu_lib().emu_mesh_laplacian, g["nodes"], bad)
    assert rc == _capi.EBADARG


def test_embedding_is_bit_exact(oracle):
    lib = oracle.load()
    tmpl = synthetic.make_template(10)
    rng = np.random.default_rng(9)
    uvn = rng.uniform(-0.7, 0.6, (400, 2))
    d = synthetic.template_surface_depth(uvn[:, 0], uvn[:, 1]) + 0.002 * rng.normal(size=400)
    pts = np.stack([uvn[:, 0] * d, uvn[:, 1] * d, d], 1).astype(np.float32)
    pts[:5] += 50.0   # far away: no facet
    rc, of, on, ob = embed_call(lib.oracle_embed_points, tmpl.nodes_rest, tmpl.facets, pts)
    rc2, ef, en, eb = embed_call(emu_lib().emu_embed_points, tmpl.nodes_rest, tmpl.facets, pts)
    nf, nn, nb = synthetic.embed_points(tmpl.nodes_rest, tmpl.facets, pts)
    assert rc == 0 and rc2 == 0
    assert np.array_equal(of, ef) and np.array_equal(on, en) and np.array_equal(ob, eb)
    assert np.array_equal(of, nf) and np.array_equal(on, nn) and np.array_equal(ob, nb)
    assert (of[:5] == -1).all() and (of >= 0).sum() > 300
    ok = of >= 0
    assert np.allclose(ob[ok].sum(1), 1.0, atol=1e-5)


def test_mappoint_recalculate(oracle):
    lib = oracle.load()
    tmpl, frames = synthetic.make_config_frames("C1", nframes=1)
    f = frames[0]
    a = np.zeros((f.n_matches, 3), np.float32)
    b = np.zeros((f.n_matches, 3), np.float32)
    args = lambda out: (tmpl.n_nodes, _capi.as_ptr(f.node_xyz, C.c_double), f.n_matches,
                        _capi.as_ptr(f.match_nodes, C.c_int32), _capi.as_ptr(f.match_bary, C.c_double),
                        _capi.as_ptr(out, C.c_float))
    assert lib.oracle_mappoints_recalculate(*args(a)) == 0
    assert emu_lib().emu_mappoints_recalculate(*args(b)) == 0
    assert np.array_equal(a, b)
