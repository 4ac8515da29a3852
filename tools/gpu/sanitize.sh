#!/bin/bash
# compute-sanitizer passes over small solves (memcheck: out-of-bounds / misaligned; racecheck: shared-memory hazards;
# synccheck: barrier misuse)
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool (sft)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/gpu/san_case.py sft 2>&1 | grep -v "^=========\s*$" | tail -8
done
for tool in memcheck racecheck; do
  echo "== $tool (nrsfm: schwarp fit, normals, shape from normals, sim3)"
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/gpu/san_case.py nrsfm 2>&1 | grep -v "^=========\s*$" | tail -6
done
echo "== memcheck (matching)"
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/gpu/san_case.py match 2>&1 | tail -5
