#!/bin/bash
# first GPU contact: parity tests + a quick timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv

timeout 600 python - <<'PY' 2>&1 | tee gpurun_out/first_timing.log
import time, numpy as np
from defslam_b200 import sft, synthetic
for cfg, nfr in [("C2", 296), ("C4", 256), ("C3", 148), ("C1", 296)]:
    tmpl, frames = synthetic.make_config_frames(cfg, nframes=8)
    frames = [frames[i % 8] for i in range(nfr)]
    T = sft.Template(tmpl)
    rb = sft.ResidentBatch(frames, template=T)
    for _ in range(2): rb.run()
    ms = [rb.run() for _ in range(5)]
    outs = rb.fetch()
    tr = np.mean([o.r.lm_trials for o in outs]); it = np.mean([o.r.lm_iterations for o in outs])
    print(cfg, "frames", nfr, rb.info(), "ms", ms, "solves/s %.0f" % (nfr / (min(ms) * 1e-3)), "iters %.1f trials %.1f" % (it, tr), flush=True)
    t = time.time(); sft.solve_batched(frames, template=T); print(" e2e call s", time.time() - t)
    rb.close(); T.close()
PY
