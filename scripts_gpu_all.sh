#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 2>&1 | tee gpurun_out/bench_n1.json | tail -1 | cut -c1-600
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_ref.json | tail -1 | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sft_lm_kernel -s 1 -c 1 -f -o gpurun_out/prof_bench_r01 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
tail -3 gpurun_out/launches_r01.csv
