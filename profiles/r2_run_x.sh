#!/bin/bash
# Round 2, GPU call X (1 x B200): K5a rank index bucketed on 1 - (1 - x)^32 instead of the top hash bits.
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback or passes or cli" > gpurun_out/r2x_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2x_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
ls gpurun_out | grep r2x
