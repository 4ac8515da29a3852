#!/bin/bash
# A/B of experiment libraries: VARIANTS="base iso ..." -> kernel ms for C2 x 2368 / C4 x 2368 / C3 x 1184 / C2 x 148
# (libdefslam_b200_<variant>.so built by __graft_entry__.build_cuda(variant=..., defines=[...]))
for v in ${VARIANTS:-base}; do
  if [ "$v" = base ]; then L=$PWD/defslam_b200/libdefslam_b200.so; else L=$PWD/defslam_b200/libdefslam_b200_$v.so; fi
  echo "=== variant $v"
  for c in "C2 2368" "C4 2368" "C3 1184" "C2 148"; do
    echo -n "$c: "; DEFSLAM_LIB=$L timeout 300 python tools/prof_run.py $c 3 2>&1 | tail -2 | tr '\n' ' '; echo
  done
done
