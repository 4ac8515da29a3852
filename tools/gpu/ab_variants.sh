#!/bin/bash
# A/B of experiment libraries: VARIANTS="base iso ..." -> kernel ms for C2 x 2368 / C4 x 2368 / C3 x 1184 and the phase cycles
for v in ${VARIANTS:-base}; do
  if [ "$v" = base ]; then L=$PWD/defslam_b200/libdefslam_b200.so; LP=$PWD/defslam_b200/libdefslam_b200_prof.so;
  else L=$PWD/defslam_b200/libdefslam_b200_$v.so; LP=$PWD/defslam_b200/libdefslam_b200_${v}_prof.so; fi
  echo "=== variant $v"
  for c in "C2 2368" "C4 2368" "C3 1184" "C2 148"; do
    echo -n "$c: "; DEFSLAM_LIB=$L timeout 300 python tools/prof_run.py $c 3 2>&1 | tail -2 | tr '\n' ' '; echo
  done
  if [ -f $LP ]; then
    for n in 148 296; do echo "-- phase cycles nprob $n"; DEFSLAM_LIB=$LP DEFSLAM_PROFILE=1 timeout 300 python tools/prof_run.py C2 $n 2 2>&1 | tail -6 | cut -c1-420; done
  fi
done
