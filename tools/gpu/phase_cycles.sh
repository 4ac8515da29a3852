#!/bin/bash
for n in ${NPROBS:-296}; do echo "== nprob $n"; DEFSLAM_LIB=$PWD/defslam_b200/libdefslam_b200_prof.so DEFSLAM_PROFILE=1 python tools/prof_run.py C2 $n 2 2>&1 | tail -6 | cut -c1-700; done
