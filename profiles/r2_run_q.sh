#!/bin/bash
# Round 2, GPU call Q (1 x B200): prune pass inside K5a (two ranks from the L1 hit count, seven ballots per group), 1024 window starts per segment.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
MM_SWEEP_PRUNE=0 MM_SWEEP_SEG=4096 timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2q_bench_noprune.json 2>> gpurun_out/r2q_bench.err
ls gpurun_out | grep r2q
