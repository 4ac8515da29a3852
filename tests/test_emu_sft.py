"""The SfT kernel sources (sft_core.h), compiled for the host as a one-thread team, against the
oracle.  This is the CPU-tier check of the kernel's arithmetic, layouts and LM control flow; the
GPU tier (test_gpu_sft.py) runs the same comparisons through the CUDA library."""
import numpy as np
import pytest

from defslam_b200 import synthetic
from tests.helpers import emu_lib, emu_normal_equations, emu_solve_batched, golden, rel_nodes

NODE_TOL = 1e-9


def _check(out, ref, f):
    assert out.r.status == 0
    assert out.r.lm_iterations == ref.r.lm_iterations
    assert out.r.lm_trials == ref.r.lm_trials
    assert rel_nodes(out.nodes, ref.nodes) < NODE_TOL
    assert np.abs(out.T_cw - ref.T_cw).max() < 1e-6
    assert np.array_equal(out.outlier[:f.n_matches], ref.outlier[:f.n_matches])
    assert out.r.n_inliers == ref.r.n_inliers
    if f.n_matches:
        assert abs(out.r.rep_error - ref.r.rep_error) < 1e-5
    k = ref.r.lm_iterations
    assert np.allclose(out.trace[:k], ref.trace[:k], rtol=1e-8, atol=1e-12)
    assert np.array_equal(out.role, ref.role)


@pytest.mark.parametrize("cfg", ["C1", "C2", "C4"])
def test_normal_equations(cfg, oracle):
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=1)
    rc, H, b, chi = emu_normal_equations(frames[0])
    assert rc == 0
    Ho, bo, chio = oracle.sft_normal_equations(frames[0])
    assert abs(chi - chio) <= 1e-12 * abs(chio)
    assert np.abs(H - Ho).max() <= 1e-12 * np.abs(Ho).max()
    assert np.abs(b - bo).max() <= 1e-12 * np.abs(bo).max()
    assert np.array_equal(H, H.T)


@pytest.mark.parametrize("cfg,nfr", [("C1", 3), ("C2", 2), ("C4", 3), ("C3", 1)])
def test_solve_matches_oracle(cfg, nfr, oracle):
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=nfr)
    rc, outs = emu_solve_batched(frames)
    assert rc == 0
    for f, o in zip(frames, outs):
        _check(o, oracle.sft_solve(f), f)


@pytest.mark.parametrize("name", ["tiny", "huber", "huber9", "C1_0", "C4_0", "mg_tiny", "mg_9"])
def test_kernel_code_matches_reference_golden(name):
    """the kernel sources against the outputs of the REFERENCE'S OWN code (tests/golden/sft_ref.npz,
    see test_oracle_sft_ref.py); 'huber*' put gross outliers on the linear branch of the Huber kernel"""
    from tests.golden.make_golden_sft import cases
    g = golden("sft_ref.npz")
    f = cases()[name]
    rc, H, b, chi = emu_normal_equations(f)
    assert rc == 0
    assert abs(chi - float(g[f"{name}.chi2"])) <= 1e-12 * abs(chi)
    assert np.abs(b - g[f"{name}.b"]).max() <= 1e-12 * np.abs(b).max()
    assert np.abs(np.diag(H) - g[f"{name}.Hdiag"]).max() <= 1e-12 * np.abs(np.diag(H)).max()
    rc, outs = emu_solve_batched([f])
    assert rc == 0
    o = outs[0]
    its, trials, inl = (int(x) for x in g[f"{name}.scalars"][:3])
    assert (o.r.lm_iterations, o.r.lm_trials, o.r.n_inliers) == (its, trials, inl)
    assert rel_nodes(o.nodes, g[f"{name}.nodes"]) < 1e-8
    assert np.array_equal(o.outlier[:f.n_matches], g[f"{name}.outlier"][:f.n_matches])
    assert np.allclose(o.trace[:its], g[f"{name}.trace"], rtol=1e-8, atol=1e-12)


def test_solve_matches_golden():
    g = golden("sft_oracle.npz")
    tmpl, frames = synthetic.make_config_frames("C1", nframes=3)
    rc, outs = emu_solve_batched(frames)
    assert rc == 0
    for i, o in enumerate(outs):
        assert np.allclose(o.nodes, g[f"C1_{i}_nodes"], rtol=0, atol=1e-9)
        assert o.r.lm_trials == int(g[f"C1_{i}_scalars"][1])


def test_partial_view_keeps_unseen_nodes_fixed(oracle):
    """Matches only in one corner: nodes outside Viewed U ring1 must not move (identity rows)."""
    tmpl = synthetic.make_template(10)
    f = synthetic.make_frame(tmpl, 400, seed=5)
    keep = (f.match_uv[:, 0] < 300) & (f.match_uv[:, 1] < 220)
    for name in ("match_nodes", "match_bary", "match_uv", "match_inv_sigma2"):
        setattr(f, name, np.ascontiguousarray(getattr(f, name)[keep]))
    rc, outs = emu_solve_batched([f])
    assert rc == 0
    ref = oracle.sft_solve(f)
    _check(outs[0], ref, f)
    fixed = ((outs[0].role >> 1) & 1) == 0
    assert fixed.sum() > 10
    assert np.array_equal(outs[0].nodes[fixed], f.node_xyz[fixed])


@pytest.mark.parametrize("kw", [dict(neighbour_layers=0), dict(reg_temp=0.0), dict(reg_lap=5000.0, reg_inex=5000.0),
                                dict(max_iterations=3), dict(n_frame_keypoints=300)])
def test_parameter_variants(kw, oracle):
    tmpl = synthetic.make_template(9)
    f = synthetic.make_frame(tmpl, 250, seed=21)
    for k, v in kw.items():
        setattr(f, k, v)
    rc, outs = emu_solve_batched([f])
    assert rc == 0
    _check(outs[0], oracle.sft_solve(f), f)


def test_moved_initial_pose_and_previous_solution_as_start(oracle):
    """Stream-style use: start from the previous frame's solution and pose."""
    tmpl = synthetic.make_template(9)
    f0 = synthetic.make_frame(tmpl, 300, seed=31)
    r0 = oracle.sft_solve(f0)
    f1 = synthetic.make_frame(tmpl, 300, seed=32)
    f1.node_xyz = np.ascontiguousarray(r0.nodes.copy())
    f1.T_cw = r0.T_cw.copy()
    rc, outs = emu_solve_batched([f1])
    assert rc == 0
    _check(outs[0], oracle.sft_solve(f1), f1)


def test_match_that_is_not_a_facet_is_rejected():
    tmpl, frames = synthetic.make_config_frames("C1", nframes=1)
    f = frames[0]
    f.match_nodes = f.match_nodes.copy()
    f.match_nodes[3] = [0, 1, tmpl.n_nodes - 1]
    rc, outs = emu_solve_batched([f])
    assert rc == -1
    assert outs[0].r.status == -1
    assert not outs[0].nodes.any()          # outputs untouched


def test_batch_results_do_not_depend_on_batch_composition():
    tmpl9 = synthetic.make_template(9)
    tmpl6 = synthetic.make_template(6)
    fa = synthetic.make_frame(tmpl9, 200, seed=41)
    fb = synthetic.make_frame(tmpl6, 60, seed=42)
    fc = synthetic.make_frame(tmpl9, 120, seed=43)
    rc, mixed = emu_solve_batched([fa, fb, fc, fa])
    assert rc == 0
    for f, o in zip([fa, fb, fc], mixed):
        rc1, alone = emu_solve_batched([f])
        assert np.array_equal(alone[0].nodes, o.nodes)
    assert np.array_equal(mixed[0].nodes, mixed[3].nodes)


def test_plan_bandwidth_of_the_regular_grid():
    import ctypes as C
    lib = emu_lib()
    for G in (6, 9, 13):
        tmpl = synthetic.make_template(G)
        h = C.c_void_p()
        assert lib.emu_template_create(C.byref(tmpl.desc()), -1, C.byref(h)) == 0
        info = (C.c_int32 * 6)()
        lib.emu_plan_info(h, info)
        assert info[0] == 3 * 2 * G + 2          # 2-ring coupling: node half-bandwidth 2G
        assert info[1] % 16 == 9 and info[1] >= info[0] + 1
        lib.emu_template_destroy(h)


def test_fallback_placements_give_identical_results(monkeypatch):
    """The planner's fallbacks of the sliding-window factorisation for large meshes (border rows, then
    x/dx, in the global workspace instead of shared memory) run the same arithmetic: bitwise identical
    solutions.  The row-owner factorisation (the default where it fits) sums in a different order:
    same LM trajectory, nodes equal to rounding."""
    tmpl, frames = synthetic.make_config_frames("C1", nframes=2)
    rc, rows = emu_solve_batched(frames)
    assert rc == 0
    monkeypatch.setenv("DEFSLAM_ROW_MODE", "0")
    rc, base = emu_solve_batched(frames)
    assert rc == 0
    for a, b in zip(rows, base):
        assert rel_nodes(a.nodes, b.nodes) < 1e-10
        assert a.r.lm_trials == b.r.lm_trials and a.r.n_inliers == b.r.n_inliers
    import ctypes as C
    lib = emu_lib()
    h = C.c_void_p()
    assert lib.emu_template_create(C.byref(tmpl.desc()), -1, C.byref(h)) == 0
    info = (C.c_int32 * 6)()
    lib.emu_plan_info(h, info)
    lib.emu_template_destroy(h)
    # shared-memory need of the default placement; just below it -> border rows move out, far below -> x/dx too
    need = info[5]
    for limit in (need - 1, need - 8 * (3 * tmpl.n_nodes) - 64):
        monkeypatch.setenv("DEFSLAM_EMU_SMEM_LIMIT", str(limit))
        rc, outs = emu_solve_batched(frames)
        assert rc == 0
        for a, b in zip(base, outs):
            assert np.array_equal(a.nodes, b.nodes)
            assert a.r.lm_trials == b.r.lm_trials and a.r.n_inliers == b.r.n_inliers
    monkeypatch.setenv("DEFSLAM_EMU_SMEM_LIMIT", "1000")
    rc, _ = emu_solve_batched(frames)
    assert rc == -4  # DEFSLAM_ETOOLARGE


@pytest.mark.parametrize("cfg,nfr", [("C1", 2), ("C2", 1), ("C4", 2), ("C3", 1)])
def test_sliding_window_factorisation_still_matches_oracle(cfg, nfr, oracle, monkeypatch):
    """the sliding-window path stays the fallback for meshes the row-owner ring does not fit (25x25)"""
    monkeypatch.setenv("DEFSLAM_ROW_MODE", "0")
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=nfr)
    rc, outs = emu_solve_batched(frames)
    assert rc == 0
    for f, o in zip(frames, outs):
        _check(o, oracle.sft_solve(f), f)


def test_stress_mesh_25x25_matches_oracle(oracle):
    """C5 (BASELINE.json stress config): 25 x 25 mesh, 2000 matches -- the placement the device uses
    for it (border rows and x/dx in the global workspace)."""
    tmpl, frames = synthetic.make_config_frames("C5", nframes=1)
    import os
    os.environ["DEFSLAM_EMU_SMEM_LIMIT"] = str((227 * 1024 - 1024) // 8)
    try:
        rc, outs = emu_solve_batched(frames)
    finally:
        del os.environ["DEFSLAM_EMU_SMEM_LIMIT"]
    assert rc == 0
    _check(outs[0], oracle.sft_solve(frames[0]), frames[0])


@pytest.mark.parametrize("G", [7, 11, 12, 14, 15, 16])
def test_row_owner_factorisation_on_every_mesh_size(G, oracle):
    """every tile count NT = 5..14 has its own instantiation of the row-owner factorisation"""
    tmpl = synthetic.make_template(G)
    f = synthetic.make_frame(tmpl, 25 * G, seed=100 + G)
    rc, outs = emu_solve_batched([f])
    assert rc == 0
    _check(outs[0], oracle.sft_solve(f), f)
