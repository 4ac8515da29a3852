#!/bin/bash
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== timings (baseline a22bbe8: C2 52.0  C4 21.7  C3 108.6 ms; 2a: 34.9 16.25 66.4)"
timeout 300 python tools/prof_run.py C2 2368 3 2>&1 | tail -1
timeout 300 python tools/prof_run.py C4 2368 3 2>&1 | tail -1
timeout 300 python tools/prof_run.py C3 1184 3 2>&1 | tail -1
NPROBS=296 bash tools/gpu/phase_cycles.sh 2>&1 | grep "cycles total\|per build\|facet sums per" | cut -c1-330
