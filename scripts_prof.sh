#!/bin/bash
# ncu evidence for the SfT LM kernel: launch list of the bench command + one full capture + phase cycles
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -c 1 -f -o gpurun_out/prof_sft python tools/prof_run.py C2 2368 1 2>&1 | tail -3
NPROBS="148 296" bash scripts_phase.sh > gpurun_out/phase_cycles.txt 2>&1
./tools/microbench > gpurun_out/microbench.txt 2>&1
tail -20 gpurun_out/phase_cycles.txt | cut -c1-300
