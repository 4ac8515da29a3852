#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python - <<'PY'
import time, numpy as np, os
from defslam_b200 import sft, synthetic
tmpl, base = synthetic.make_config_frames("C2", nframes=64)
frames = [base[i % 64] for i in range(2368)]
T = sft.Template(tmpl)
hb = sft.HostBatch(frames, template=T)
for _ in range(2): hb.solve()
for tag in ("pipelined",):
    ts = []
    for _ in range(5):
        t = time.perf_counter(); hb.solve(); ts.append(time.perf_counter() - t)
    print(tag, "e2e ms", np.round(np.array(ts) * 1e3, 2), "solves/s %.0f" % (2368 / np.median(ts)))
rb = sft.ResidentBatch(frames, template=T)
print("resident ms", [round(rb.run(), 2) for _ in range(3)])
PY
