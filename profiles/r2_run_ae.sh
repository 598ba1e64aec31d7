#!/bin/bash
# Round 2, GPU call AE (1 x B200): caching device allocator -- all GPU tests, config-5 slice three times, config 4 twice, the bench line.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ae_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ae_tests.log)
for i in 1 2 3; do timeout 600 python bench.py --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ae_c5_$i.json 2>> gpurun_out/r2ae.err; done
MM_ALLOC_CACHE_GB=0 timeout 600 python bench.py --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ae_c5_nocache.json 2>> gpurun_out/r2ae.err
for i in 1 2; do timeout 600 python bench.py --workload config4 --reads 100000 > gpurun_out/r2ae_c4_$i.json 2>> gpurun_out/r2ae.err; done
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ae_bench.json 2>> gpurun_out/r2ae.err
ls gpurun_out | grep r2ae
