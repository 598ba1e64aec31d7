#!/bin/bash
# Round 2, GPU call U (1 x B200): cooperative first-window build in K5b; staging default = DMA (zerocopy leg for comparison).
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback or passes or cli or staged" > gpurun_out/r2u_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2u_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
MM_STAGE=zerocopy timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2u_bench_zerocopy.json 2>> gpurun_out/r2u_bench.err
ls gpurun_out | grep r2u
