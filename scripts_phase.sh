#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -3
for n in 1 296; do echo "== nprob $n"; DEFSLAM_LIB=$PWD/defslam_b200/libdefslam_b200_prof.so DEFSLAM_PROFILE=1 python tools/prof_run.py C2 $n 2 2>&1 | tail -2 | cut -c1-400; done
bash scripts_first_gpu.sh 2>&1 | grep -E "solves/s"
