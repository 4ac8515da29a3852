#!/bin/bash
# ncu evidence: launch list of the bench command + full captures of the LM kernel and the NRSfM kernels + phase cycles
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -c 1 -f -o gpurun_out/prof_sft python tools/prof_run.py C2 2368 1 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'schwarp_fit_kernel|sfn_solve_kernel|normals_kernel' -c 3 -f -o gpurun_out/prof_nrsfm python tools/nrsfm_prof.py 2>&1 | tail -2
NPROBS="148 296" bash tools/gpu/phase_cycles.sh > gpurun_out/phase_cycles.txt 2>&1
tail -8 gpurun_out/phase_cycles.txt | cut -c1-200
