#!/bin/bash
# Round 2, GPU call V (1 x B200): pruned K5b, ring depth x warps per SM, segment length.
set -x
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2v_$name.json 2>> gpurun_out/r2v.err; }
run ring8_w12 MM_SWEEP_RING=8
run ring4_w12 MM_SWEEP_RING=4 MM_SWEEP_WARPS=12
run ring4_w16 MM_SWEEP_RING=4 MM_SWEEP_WARPS=16
run ring2_w16 MM_SWEEP_RING=2 MM_SWEEP_WARPS=16
run ring2_w12 MM_SWEEP_RING=2 MM_SWEEP_WARPS=12
run seg768 MM_SWEEP_SEG=768
run seg1536 MM_SWEEP_SEG=1536
ls gpurun_out | grep r2v
