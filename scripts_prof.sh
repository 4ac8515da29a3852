#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'schwarp_fit_kernel|sfn_solve_kernel' -c 2 -f -o gpurun_out/prof_nrsfm_r01 python tools/nrsfm_prof.py 2>&1 | tail -5
