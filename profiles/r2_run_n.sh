#!/bin/bash
# Round 2, GPU call N (1 x B200): window pruning between K5a and K5b (l2_prune_warp_kernel) -- parity tests, bench with / without pruning.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2n_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
MM_SWEEP_PRUNE=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2n_bench_noprune.json 2>> gpurun_out/r2n_bench.err
ls gpurun_out | grep r2n
