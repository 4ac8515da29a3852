#!/bin/bash
echo "== 2 CTAs/SM (product)"; timeout 300 python tools/prof_run.py C2 2368 3 2>&1 | tail -1
echo "== forced global E/x, still 2 CTAs/SM (cost of the placement alone)"
DEFSLAM_SMEM_LIMIT=75000 timeout 300 python tools/prof_run.py C2 2368 3 2>&1 | tail -1
echo "== 3 CTAs/SM variant (85 regs, E and x/dx in global)"
DEFSLAM_LIB=$PWD/defslam_b200/libdefslam_b200_c3.so DEFSLAM_SMEM_LIMIT=75000 timeout 300 python tools/prof_run.py C2 2664 3 2>&1 | tail -1
DEFSLAM_LIB=$PWD/defslam_b200/libdefslam_b200_c3.so DEFSLAM_SMEM_LIMIT=75000 timeout 300 python -c "
from defslam_b200 import sft, synthetic
tmpl, fr = synthetic.make_config_frames('C2', nframes=4)
rb = sft.ResidentBatch([fr[i%4] for i in range(2664)], template=sft.Template(tmpl)); print(rb.info())"
