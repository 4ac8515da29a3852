"""One launch of each NRSfM stage kernel on the bench-shaped workload (for ncu): 592 Schwarp fits, ~480 k map-point
normals, 296 shape-from-normals keyframes -- the same units bench.py times."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

os.environ["DEFSLAM_NORMALS_CHUNKS"] = "1"  # the normals batch in ONE launch (the call pipelines it in chunks otherwise)

wl = bench.nrsfm_workload()
api = wl["api"]
api.schwarp_prepare(wl["pairs"])()
api.normals_prepare(wl["normals"])()
api.sfn_prepare(wl["keyframes"])()
print("pairs", len(wl["pairs"]), "points", wl["normals"].n, "keyframes", len(wl["keyframes"]))
