#!/bin/bash
timeout 600 python -m pytest tests/test_matching.py tests/test_newpoints.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python - <<'PY'
import time, numpy as np
from defslam_b200 import matching, _capi
lib = _capi.load()
c = matching.make_case(seed=1)
for _ in range(3): matching.search_by_projection(c)
t = time.perf_counter(); n = 50
for _ in range(n): m, nm = matching.search_by_projection(c)
dt = (time.perf_counter() - t) / n
print("projection search: %d map points x %d keypoints, %d matches, %.3f ms per call (host buffers)" % (len(c.last_state), len(c.cur_octave), nm, dt * 1e3))
PY
