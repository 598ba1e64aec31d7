#!/bin/bash
# Round 2, GPU call AS (1 x B200): compute-sanitizer (memcheck, racecheck, synccheck) over the map path on small workloads (golden, config 1, fallbacks).
set -x
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map_golden or ragged or fuzz" > gpurun_out/r2as_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2as_$tool.log
done
ls gpurun_out | grep r2as
