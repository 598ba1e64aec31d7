#!/bin/bash
# Round 2, GPU call AI (1 x B200): 32-byte slots, inline ids used / not used by the fused K4 kernel; table at load 0.25.
set -x
mkdir -p gpurun_out
MM_K4_INLINE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ai_inline1.json 2>> gpurun_out/r2ai.err
MM_K4_INLINE=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ai_inline0.json 2>> gpurun_out/r2ai.err
MM_K4_INLINE=1 MM_L1_CTAS=6 MM_L1_CACHE=6144 timeout 300 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ai_inline1_c6.json 2>> gpurun_out/r2ai.err
ls gpurun_out | grep r2ai
