#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -c 1 -f -o gpurun_out/prof_sft python tools/prof_run.py C2 2368 1 2>&1 | tail -3
