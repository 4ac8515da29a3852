#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-6000; tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_ref.json | cut -c1-1200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
