#!/bin/bash
# Round 2, GPU call AT (1 x B200): K5a's bucket table: the run of empty buckets above a sketch's largest hash filled by the whole CTA.
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map or variants or fallback or passes or cli or fuzz" > gpurun_out/r2at_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2at_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2at_bench.json 2> gpurun_out/r2at_bench.err
ls gpurun_out | grep r2at
