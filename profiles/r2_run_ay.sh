#!/bin/bash
# Round 2, GPU call AY (1 x B200): last check of the committed tree -- all GPU tests, smoke, a short bench line.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ay_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ay_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ay_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2ay_smoke.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ay_bench.json 2> gpurun_out/r2ay_bench.err
ls gpurun_out | grep r2ay
