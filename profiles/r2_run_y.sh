#!/bin/bash
# Round 2, GPU call Y (1 x B200): config 4 (200 candidates per read) with / without the window pruning, standalone and as the extra leg of the default run.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload config4 --reads 100000 > gpurun_out/r2y_c4_prune.json 2> gpurun_out/r2y.err
MM_SWEEP_PRUNE=0 timeout 600 python bench.py --workload config4 --reads 100000 > gpurun_out/r2y_c4_noprune.json 2>> gpurun_out/r2y.err
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2y_bench_prune.json 2>> gpurun_out/r2y.err
ls gpurun_out | grep r2y
