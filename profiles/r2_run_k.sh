#!/bin/bash
# Round 2, GPU call K2 (1 x B200): fused K4 with reads handed out dynamically -- CTAs per SM sweep.
set -x
mkdir -p gpurun_out
for v in "512 8192 3" "512 8192 4" "512 8192 5" "512 8192 6" "1024 8192 5" "512 12288 5" "512 6144 7"; do
  set -- $v
  MM_L1_CTAS=$3 MM_L1_CHUNK=$1 MM_L1_CACHE=$2 timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2k2_chunk$1_cache$2_ctas$3.json 2>> gpurun_out/r2k2.err
done
ls gpurun_out | grep r2k2
