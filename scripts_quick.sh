#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_sft.py -x -q -m gpu 2>&1 | tail -2
timeout 300 python tools/prof_run.py C2 2368 4 2>&1 | tail -3
timeout 300 python tools/prof_run.py C4 2368 3 2>&1 | tail -2
timeout 300 python tools/prof_run.py C3 1184 3 2>&1 | tail -2
