#!/bin/bash
# Round 2, GPU call AM (1 x B200): per-read bitonic sort of the L1 survivors in shared memory instead of the device-wide radix sort.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2am_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2am_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2am_bench.json 2> gpurun_out/r2am_bench.err
MM_L1_SEGSORT=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2am_bench_radix.json 2>> gpurun_out/r2am_bench.err
ls gpurun_out | grep r2am
