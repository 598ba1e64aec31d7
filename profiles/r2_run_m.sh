#!/bin/bash
# Round 2, GPU call M (1 x B200): K3 radix bits per pass (4 / 5 / 6), e2e stage breakdown.
set -x
mkdir -p gpurun_out
for v in 4 5 6; do
  MM_K3_RADIX=$v timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2m_k3radix$v.json 2>> gpurun_out/r2m.err
done
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-extras --e2e-breakdown > gpurun_out/r2m_breakdown.json 2> gpurun_out/r2m_breakdown.err
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map_golden or config1 or ragged or fallback" > gpurun_out/r2m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_tests.log)
ls gpurun_out | grep r2m
