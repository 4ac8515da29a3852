#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-3500; tail -5 gpurun_out/bench_n1.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
