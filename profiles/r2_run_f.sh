#!/bin/bash
# Round 2, GPU call F (1 x B200): what bounds K5b?  resident warps per SM (MM_SWEEP_WARPS) x event-ring depth (MM_SWEEP_RING).
set -x
mkdir -p gpurun_out
for v in "8 8" "8 12" "4 12" "4 16" "2 12" "2 16"; do
  set -- $v
  MM_SWEEP_RING=$1 MM_SWEEP_WARPS=$2 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2f_ring$1_warps$2.json 2>> gpurun_out/r2f.err
done
ls gpurun_out | grep r2f
