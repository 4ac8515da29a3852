"""Device time of the Schwarp fit kernel with parts of the fit switched off (diagnostics): the differences
give the cost of the initialisation solve and of one trust-region iteration."""
import sys, dataclasses
sys.path.insert(0, ".")
from defslam_b200 import nrsfm, _capi

lib = _capi.load()
api = nrsfm.Api(lib, "defslam_")
wins = [nrsfm.make_window(100 + i, n_keypoints=1200, n_views=4) for i in range(8)]
cases = [c for w in wins for c in nrsfm.schwarp_cases(w)]
fits = api.schwarp_fit_batched(cases)
batch = cases * (592 // len(cases))
for init, iters in ((1, 3), (1, 0), (0, 0), (0, 1), (0, 3)):
    bb = []
    for c, f in zip(batch, fits * 100):
        c2 = dataclasses.replace(c, initialize=init, max_iterations=iters)
        if not init:
            c2.x0 = f.x
        bb.append(c2)
    for _ in range(2):
        outs = api.schwarp_fit_batched(bb)
    print("initialize=%d max_iterations=%d: kernel %.3f ms for %d fits" % (
        init, iters, lib.defslam_last_kernel_ms(), len(bb)), flush=True)

import os
if "nprof" in os.environ.get("DEFSLAM_LIB", ""):
    import numpy as np
    print("diagnostic build: cycles in the LM solves %.0f, LM loop + DiffProp %.0f (mean per fit, last configuration)" % (
        np.mean([o.d.cost_initial for o in outs]), np.mean([o.d.cost_final for o in outs])))
