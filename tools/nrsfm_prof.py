"""One launch of each NRSfM stage kernel on a bench-shaped workload (for ncu)."""
import sys
sys.path.insert(0, ".")
from defslam_b200 import nrsfm
api = nrsfm.Api()
wins = [nrsfm.make_window(100 + i, n_keypoints=1200, n_views=4) for i in range(4)]
cases = [c for w in wins for c in nrsfm.schwarp_cases(w)]
fits = api.schwarp_fit_batched(cases * 37)   # 592 pairs
ncs = [nrsfm.normals_case(w, fits[4 * i:4 * i + 4]) for i, w in enumerate(wins)]
nouts = [api.normals(nc) for nc in ncs]
scs = [nrsfm.sfn_case(w, no) for w, no in zip(wins, nouts)]
api.sfn_solve_batched(scs * 74)              # 296 keyframes
