#!/bin/bash
# Round 2, GPU call AX (1 x B200): per-read survivor sort, 512 keys per warp (six CTAs per SM) instead of 1024.  (variant not kept; the code it measured is described in DESIGN.md section 4 and was reverted)
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback" > gpurun_out/r2ax_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ax_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ax_bench.json 2> gpurun_out/r2ax_bench.err
