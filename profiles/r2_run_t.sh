#!/bin/bash
# Round 2, GPU call T (1 x B200): prune decision in its own warp-per-candidate kernel; e2e with DMA staging (MM_STAGE=copy selected it then; it is the default now).  (variant not kept; the code it measured is described in DESIGN.md section 4 and was reverted)
set -x
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback or passes or cli" > gpurun_out/r2t_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2t_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
MM_STAGE=copy timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2t_bench_dma.json 2>> gpurun_out/r2t_bench.err
