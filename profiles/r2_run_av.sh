#!/bin/bash
# Round 2, GPU call AV (1 x B200): the committed tree -- all GPU tests, smoke, the bench line as the driver runs it, the launch list.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2av_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2av_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2av_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2av_smoke.log)
timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2av_launches.csv \
  python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2av_launch.log 2>&1
ls -la gpurun_out | grep r2av
