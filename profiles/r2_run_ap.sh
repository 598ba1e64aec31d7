#!/bin/bash
# Round 2, GPU call AP (1 x B200): where the e2e step differs from the device-resident step (DMA staging), per-stage.
set -x
mkdir -p gpurun_out
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --e2e-breakdown > gpurun_out/r2ap_breakdown.json 2> gpurun_out/r2ap_breakdown.err
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "variants" > gpurun_out/r2ap_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ap_tests.log)
ls gpurun_out | grep r2ap
