#!/bin/bash
# Round 2, GPU call AD (1 x B200): config-5 slice, where the time of a chunk goes (fused / two-kernel K4, pruning on / off).
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ad_c5.json 2> gpurun_out/r2ad.err
MM_L1_FUSED=0 timeout 600 python bench.py --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ad_c5_twokernels.json 2>> gpurun_out/r2ad.err
MM_SWEEP_PRUNE=0 timeout 600 python bench.py --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ad_c5_noprune.json 2>> gpurun_out/r2ad.err
ls gpurun_out | grep r2ad
