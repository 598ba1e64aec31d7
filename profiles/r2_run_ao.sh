#!/bin/bash
# Round 2, GPU call AO (1 x B200): candidate regions by one warp per read over its sorted hits (one kernel + a gather instead of four kernels and two scans).
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ao_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ao_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ao_bench.json 2> gpurun_out/r2ao_bench.err
MM_L1_CAND=legacy timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ao_bench_legacy.json 2>> gpurun_out/r2ao_bench.err
ls gpurun_out | grep r2ao
