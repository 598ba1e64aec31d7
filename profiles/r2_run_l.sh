#!/bin/bash
# Round 2, GPU call L (1 x B200): fused K4 with dynamic read hand-out (defaults), K3 with dynamic hand-out -- parity tests + bench.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err
MM_L1_FUSED=0 timeout 600 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2l_bench_twokernels.json 2>> gpurun_out/r2l_bench.err
ls gpurun_out | grep r2l
