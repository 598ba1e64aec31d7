#!/bin/bash
# Round 2, GPU call G (1 x B200): K5b window skipping -- parity tests, bench with and without it, ncu of K5a / K5b.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2g_tests.log)
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
MM_SWEEP_SKIP=0 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2g_bench_noskip.json 2>> gpurun_out/r2g_var.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem' \
  -o gpurun_out/r2g_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
ls gpurun_out | grep r2g
