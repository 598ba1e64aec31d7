#!/bin/bash
# Round 2, GPU call J (1 x B200): K4 as one kernel (probe + contig filter + gather) -- parity tests, then A/B against the two-kernel route.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log)
timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_fused.json 2> gpurun_out/r2j_bench.err
MM_L1_FUSED=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_twokernels.json 2>> gpurun_out/r2j_bench.err
MM_L1_CACHE=8192 timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_fused_cache8k.json 2>> gpurun_out/r2j_bench.err
MM_L1_CACHE=4096 timeout 600 python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2j_bench_fused_cache4k.json 2>> gpurun_out/r2j_bench.err
ls gpurun_out | grep r2j
