#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -3
DEFSLAM_PROFILE=1 python tools/prof_run.py C2 296 2
DEFSLAM_PROFILE=1 python tools/prof_run.py C2 1 2
DEFSLAM_PROFILE=1 python tools/prof_run.py C4 256 2
