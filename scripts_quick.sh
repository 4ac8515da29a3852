#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -2
bash scripts_first_gpu.sh 2>&1 | grep -E "solves/s" | cut -c1-60,200-420
