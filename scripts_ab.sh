#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -2
for lib in libdefslam_b200_old.so libdefslam_b200.so; do
  echo "== $lib"
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 python tools/prof_run.py C2 2368 3 2>&1 | tail -1
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 python tools/prof_run.py C4 2368 3 2>&1 | tail -1
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 python tools/prof_run.py C3 1184 3 2>&1 | tail -1
done
bash scripts_phase.sh
