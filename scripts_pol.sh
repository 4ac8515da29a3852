#!/bin/bash
for lib in libdefslam_b200.so libdefslam_b200_p1.so libdefslam_b200_p2.so; do
  echo "== $lib"
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 python tools/prof_run.py C2 2368 3 2>&1 | tail -1
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 python tools/prof_run.py C3 1184 3 2>&1 | tail -1
  DEFSLAM_LIB=$PWD/defslam_b200/$lib timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:sft_lm -c 1 python tools/prof_run.py C2 2368 1 2>&1 | grep "dram__bytes"
done
