#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
bash scripts_first_gpu.sh 2>&1 | grep -E "solves/s"
timeout 600 python tools/nrsfm_timing.py 8 > gpurun_out/nrsfm_timing.log 2>&1; tail -8 gpurun_out/nrsfm_timing.log
