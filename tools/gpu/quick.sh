#!/bin/bash
timeout 600 python -m pytest tests/test_matching.py tests/test_adapter_cpp.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python - <<'PY'
import time, numpy as np
from defslam_b200 import matching, _capi
from oracle import oracle_py
lib = _capi.load(); olib = oracle_py.load()
for name, c, fn in (("projection", matching.make_case(1), matching.search_by_projection), ("schwarp", matching.make_warp_case(1), matching.search_by_schwarp)):
    for _ in range(3): fn(c)
    t = time.perf_counter(); n = 100
    for _ in range(n): m, nm = fn(c)
    dt = (time.perf_counter() - t) / n
    t = time.perf_counter()
    for _ in range(n): fn(c, olib, "oracle_")
    dto = (time.perf_counter() - t) / n
    print("%s: GPU %.3f ms per call (host buffers), oracle %.3f ms, %d matches" % (name, dt * 1e3, dto * 1e3, nm))
PY
