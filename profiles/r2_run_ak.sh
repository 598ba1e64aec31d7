#!/bin/bash
# Round 2, GPU call AK (1 x B200): prune pass up to 256 / 512 groups per candidate, twice each (K5b's run-to-run spread).
set -x
mkdir -p gpurun_out
for i in 1 2; do for g in 256 512; do
  MM_PRUNE_GMAX=$g timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ak_g${g}_$i.json 2>> gpurun_out/r2ak.err
done; done
ls gpurun_out | grep r2ak
