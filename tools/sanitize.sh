#!/bin/bash
# compute-sanitizer over the small fixtures: memcheck + racecheck of every kernel of one train of
# `tiny` (dense tables, all bins that exist at that size) and `small` (dense + hashed tables, CTA-owned rows, prunes).
# usage (on the GPU box): bash tools/sanitize.sh <tag>   -> gpurun_out/<tag>_{memcheck,racecheck}_{tiny,small}.log
TAG=$1
for wl in tiny small; do
  QC_ITERS=1 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/quick_check.py $wl > gpurun_out/${TAG}_memcheck_${wl}.log 2>&1
  QC_ITERS=1 timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/quick_check.py $wl > gpurun_out/${TAG}_racecheck_${wl}.log 2>&1
done
