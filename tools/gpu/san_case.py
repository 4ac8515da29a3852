"""Small workloads for compute-sanitizer (tools/gpu/sanitize.sh)."""
import os
import sys

sys.path.insert(0, os.getcwd())
from defslam_b200 import matching, sft, synthetic  # noqa: E402

which = sys.argv[1]
if which == "sft":
    for cfg, n in (("C1", 2), ("C2", 1)):
        tmpl, frames = synthetic.make_config_frames(cfg, nframes=n)
        for f in frames:
            f.max_iterations = 3
        outs = sft.solve_batched(frames)
        print(cfg, [o.r.lm_trials for o in outs])
elif which == "nrsfm":
    from defslam_b200 import nrsfm
    api = nrsfm.Api()
    win = nrsfm.make_window(3, n_keypoints=160, n_views=2)
    fits = api.schwarp_fit_batched(nrsfm.schwarp_cases(win))
    nout = api.normals(nrsfm.normals_case(win, fits))
    os.environ["DEFSLAM_NORMALS_CHUNKS"] = "3"   # the pipelined path (chunks over two side streams) on the same small case
    nout3 = api.normals(nrsfm.normals_case(win, fits))
    del os.environ["DEFSLAM_NORMALS_CHUNKS"]
    assert (nout3.status == nout.status).all() and (nout3.k == nout.k).all()
    ctrl, xyz = api.sfn_solve(nrsfm.sfn_case(win, nout))
    r = api.sim3_register([nrsfm.sim3_case(2, n=120)])
    print("nrsfm", int((nout.status == 1).sum()), float(ctrl[0]), r[0]["inliers"])
else:
    c = matching.make_case(1, n_last=300, n_clutter=200)
    print(matching.search_by_projection(c)[1])
    w = matching.make_warp_case(1, n1=300, n_clutter=200)
    print(matching.search_by_schwarp(w)[1])
