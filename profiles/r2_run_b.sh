#!/bin/bash
# Round 2, GPU call B (1 x B200): pipelined K5b + table load 0.25 + threaded host; config 4 at full size; config 5 slice.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_tests.log)
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
MM_TABLE_MULT=2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_mult2.json 2>> gpurun_out/r2b_var.err
MM_SWEEP_RING=4 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2b_bench_ring4.json 2>> gpurun_out/r2b_var.err
timeout 600 python bench.py --workload config5-tiny --steps 3 --warmup 1 > gpurun_out/r2b_config5tiny.json 2> gpurun_out/r2b_config5tiny.err
timeout 900 python bench.py --workload config5-slice --steps 3 --warmup 1 > gpurun_out/r2b_config5.json 2> gpurun_out/r2b_config5.err
timeout 900 python bench.py --workload config4 > gpurun_out/r2b_config4.json 2> gpurun_out/r2b_config4.err
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l1_probe_tma|l1_filter_gather16' \
  -o gpurun_out/r2b_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2b_ncu.log 2>&1
ls -la gpurun_out | grep r2b
