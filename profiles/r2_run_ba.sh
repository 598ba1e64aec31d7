#!/bin/bash
# Round 2, GPU call BA (1 x B200): ncu --set full (raw + source pages) of the committed tree's large kernels, incl. the per-read sort and the candidate kernel.
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem|l2_prune_warp|l1_probe_filter|sketch_blockmin_kernel|l1_sort_segments_warp|l1_candidates_warp' \
  -o /tmp/r2ba_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2ba_ncu.log 2>&1
ncu -i /tmp/r2ba_full.ncu-rep --page raw --csv > gpurun_out/r2ba_full_raw.csv 2>/dev/null
ncu -i /tmp/r2ba_full.ncu-rep --page source --csv > gpurun_out/r2ba_source.csv 2>/dev/null
ls -la gpurun_out | grep r2ba
