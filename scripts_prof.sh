#!/bin/bash
# ncu evidence for the SfT LM kernel: launch list of the bench command + one full capture + phase cycles
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -c 1 -f -o gpurun_out/prof_sft python tools/prof_run.py C2 2368 1 2>&1 | tail -3
bash scripts_phase.sh
