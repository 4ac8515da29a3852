#!/bin/bash
# ncu evidence (round tag R, default r02): launch list of the bench command, full captures of the SfT LM kernel and of
# the three NRSfM kernels at bench-size grids, phase cycles of the profile build.  Outputs under gpurun_out/.
R=${R:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -c 1 -f -o gpurun_out/prof_sft_$R python tools/prof_run.py C2 2368 1 2>&1 | tail -2
# the last launch of each NRSfM kernel is the bench-size one (the workload builder runs small launches first)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'schwarp_fit_kernel' --launch-skip 1 -c 1 -f -o gpurun_out/prof_schwarp_$R python tools/nrsfm_prof.py 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'normals_kernel' --launch-skip 8 -c 1 -f -o gpurun_out/prof_normals_$R python tools/nrsfm_prof.py 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sfn_solve_kernel' -c 1 -f -o gpurun_out/prof_sfn_$R python tools/nrsfm_prof.py 2>&1 | tail -2
NPROBS="148 296" bash tools/gpu/phase_cycles.sh > gpurun_out/phase_cycles_$R.txt 2>&1
tail -8 gpurun_out/phase_cycles_$R.txt | cut -c1-200
