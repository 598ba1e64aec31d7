#!/bin/bash
# Round 2, GPU call W (1 x B200): K5a rank index with 2048 / 4096 / 8192 buckets.
set -x
mkdir -p gpurun_out
for v in 11 12 13; do
  MM_CLS_BITS=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2w_bits$v.json 2>> gpurun_out/r2w.err
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map_golden or config1 or at_scale" > gpurun_out/r2w_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_tests.log)
ls gpurun_out | grep r2w
