#!/bin/bash
# Round 2, GPU call S (1 x B200): ncu --set full with source counters of K5a with the prune pass.
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_classify_smem' \
  -o /tmp/r2s_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2s_ncu.log 2>&1
ncu -i /tmp/r2s_full.ncu-rep --page raw --csv > gpurun_out/r2s_full_raw.csv 2>/dev/null
ncu -i /tmp/r2s_full.ncu-rep --page source --csv > gpurun_out/r2s_source.csv 2>/dev/null
ls -la gpurun_out | grep r2s
