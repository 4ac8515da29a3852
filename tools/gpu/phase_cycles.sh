#!/bin/bash
# per-phase cycle counters of the profile build (libdefslam_b200_prof.so) for the C2 batch at NPROBS frames
for n in ${NPROBS:-296}; do echo "== nprob $n"; DEFSLAM_LIB=$PWD/defslam_b200/libdefslam_b200_prof.so DEFSLAM_PROFILE=1 python tools/prof_run.py ${CFG:-C2} $n 2 2>&1 | tail -8 | cut -c1-700; done
