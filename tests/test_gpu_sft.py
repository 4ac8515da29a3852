"""GPU parity tests of the SfT solve: CUDA path (through the C ABI) vs the CPU oracle."""
import numpy as np
import pytest

from defslam_b200 import sft, synthetic

pytestmark = pytest.mark.gpu

NODE_TOL = 1e-8   # relative to RMS node norm; north_star allows 1e-4


def _rel_nodes(a, b):
    return np.abs(a - b).max() / np.sqrt((b ** 2).sum(1).mean())


@pytest.mark.parametrize("cfg", ["C1", "C2", "C4", "C5"])
def test_normal_equations_match_oracle(cfg, oracle, cuda_lib):
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=1)
    H, b, chi = sft.normal_equations(frames[0])
    Ho, bo, chio = oracle.sft_normal_equations(frames[0])
    assert abs(chi - chio) <= 1e-12 * abs(chio)
    assert np.abs(H - Ho).max() <= 1e-12 * np.abs(Ho).max()
    assert np.abs(b - bo).max() <= 1e-12 * np.abs(bo).max()


@pytest.mark.parametrize("cfg,nframes", [("C1", 3), ("C2", 2), ("C4", 4), ("C3", 1), ("C5", 2)])
def test_solve_matches_oracle(cfg, nframes, oracle, cuda_lib):
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=nframes)
    outs = sft.solve_batched(frames)
    for f, o in zip(frames, outs):
        ref = oracle.sft_solve(f)
        assert o.r.status == 0
        assert o.r.lm_iterations == ref.r.lm_iterations
        assert o.r.lm_trials == ref.r.lm_trials
        assert _rel_nodes(o.nodes, ref.nodes) < NODE_TOL
        assert np.abs(o.T_cw - ref.T_cw).max() < 1e-6
        assert (o.outlier[:f.n_matches] == ref.outlier[:f.n_matches]).all()
        assert o.r.n_inliers == ref.r.n_inliers
        assert abs(o.r.rep_error - ref.r.rep_error) < 1e-5
        k = ref.r.lm_iterations
        assert np.allclose(o.trace[:k], ref.trace[:k], rtol=1e-8, atol=1e-12)
        assert (o.role == ref.role).all()


def test_template_handle_and_resident_batch(oracle, cuda_lib):
    tmpl, frames = synthetic.make_config_frames("C4", nframes=8)
    T = sft.Template(tmpl)
    info = T.info()
    assert info["bandwidth"] == 3 * 2 * tmpl.G + 2
    a = sft.solve_batched(frames, template=T)
    rb = sft.ResidentBatch(frames, template=T)
    ms = rb.run()
    assert ms > 0
    b = rb.fetch()
    rb.run()
    c = rb.fetch()
    for x, y, z in zip(a, b, c):
        # same kernel, same inputs: bitwise reproducible
        assert np.array_equal(x.nodes, y.nodes) and np.array_equal(y.nodes, z.nodes)
        assert x.r.lm_trials == y.r.lm_trials == z.r.lm_trials
    ref = oracle.sft_solve(frames[3])
    assert _rel_nodes(a[3].nodes, ref.nodes) < NODE_TOL
    rb.close()
    T.close()


def test_bad_match_is_rejected(cuda_lib):
    tmpl, frames = synthetic.make_config_frames("C1", nframes=1)
    f = frames[0]
    f.match_nodes = f.match_nodes.copy()
    f.match_nodes[5] = [0, 1, tmpl.n_nodes - 1]  # not a facet
    with pytest.raises(sft.DefslamError) as e:
        sft.solve_batched([f])
    assert e.value.rc == -1


def test_pipelined_host_path_equals_single_launch(cuda_lib, monkeypatch):
    """large host batches are cut into chunks that overlap marshalling with the kernel: same results"""
    tmpl, base = synthetic.make_config_frames("C1", nframes=8)
    frames = [base[i % 8] for i in range(3 * 4 * 148 + 37)]
    T = sft.Template(tmpl)
    piped = sft.solve_batched(frames, template=T)
    monkeypatch.setenv("DEFSLAM_NO_PIPELINE", "1")
    single = sft.solve_batched(frames, template=T)
    for a, b in zip(piped, single):
        assert a.r.status == 0 and b.r.status == 0
        assert np.array_equal(a.nodes, b.nodes)
        assert a.r.lm_trials == b.r.lm_trials and a.r.n_inliers == b.r.n_inliers
        assert np.array_equal(a.outlier, b.outlier)
    T.close()


def test_ragged_batch_of_different_templates(oracle, cuda_lib):
    """one launch over frames of different meshes and match counts (the planner sizes shared memory and
    the workspace for the largest; every CTA re-derives its layout per frame), plus a frame without matches"""
    t9, t6, t13 = synthetic.make_template(9), synthetic.make_template(6), synthetic.make_template(13)
    frames = [synthetic.make_frame(t9, 200, seed=41), synthetic.make_frame(t6, 60, seed=42),
              synthetic.make_frame(t13, 700, seed=44), synthetic.make_frame(t9, 120, seed=43)]
    empty = synthetic.make_frame(t9, 50, seed=45)
    for name in ("match_nodes", "match_bary", "match_uv", "match_inv_sigma2"):
        setattr(empty, name, np.ascontiguousarray(getattr(empty, name)[:0]))
    frames.append(empty)
    frames = frames * 3
    outs = sft.solve_batched(frames)
    for f, o in zip(frames, outs):
        ref = oracle.sft_solve(f)
        assert o.r.status == 0
        assert o.r.lm_iterations == ref.r.lm_iterations and o.r.lm_trials == ref.r.lm_trials
        assert _rel_nodes(o.nodes, ref.nodes) < NODE_TOL
        assert o.r.n_inliers == ref.r.n_inliers
    assert outs[4].r.n_inliers == 0 and np.array_equal(outs[4].nodes, outs[9].nodes)
