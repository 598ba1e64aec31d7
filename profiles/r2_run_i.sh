#!/bin/bash
# Round 2, GPU call I (1 x B200): EM round kernel with warp-owned taxon sums + four loads in flight; CLI phase timing.
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_em or mapq or pipeline or multi_batch or save_load" > gpurun_out/r2i_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2i_tests.log)
timeout 900 python bench.py --workload config4 --reads 100000 > gpurun_out/r2i_config4_100k.json 2> gpurun_out/r2i_config4_100k.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
ls gpurun_out | grep r2i
