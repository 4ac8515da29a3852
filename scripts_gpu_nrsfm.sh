#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/nrsfm_timing.py 8 > gpurun_out/nrsfm_timing.log 2>&1; tail -12 gpurun_out/nrsfm_timing.log
