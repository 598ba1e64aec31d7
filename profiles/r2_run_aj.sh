#!/bin/bash
# Round 2, GPU call AJ (1 x B200): prune pass for spans of up to 16384 elements (512 groups), two-pass scan in the prune kernel.
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback or passes or cli or fuzz" > gpurun_out/r2aj_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2aj_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err
ls gpurun_out | grep r2aj
