#!/bin/bash
# Round 2, GPU call AN (1 x B200): the per-read survivor sort as an all-ascending bitonic network without padding.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2an_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2an_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err
MM_L1_SEGSORT=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2an_bench_radix.json 2>> gpurun_out/r2an_bench.err
ls gpurun_out | grep r2am
