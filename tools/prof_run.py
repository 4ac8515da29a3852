"""Small driver for profiling: one resident batch of a config, N kernel runs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from defslam_b200 import sft, synthetic  # noqa: E402

cfg = sys.argv[1]
nfr = int(sys.argv[2])
reps = int(sys.argv[3])
tmpl, frames = synthetic.make_config_frames(cfg, nframes=4)
frames = [frames[i % 4] for i in range(nfr)]
T = sft.Template(tmpl)
rb = sft.ResidentBatch(frames, template=T)
for _ in range(reps):
    print(rb.run())
