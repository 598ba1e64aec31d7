#!/bin/bash
# Round 2, GPU call AF (8 x B200): final tree, N = 8 bench line without the extra legs, DMA staging (default) against zero-copy staging.
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2af_bench_n8.json 2> gpurun_out/r2af_bench_n8.err
MM_STAGE=zerocopy timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2af_bench_n8_zerocopy.json 2> gpurun_out/r2af_bench_n8_zerocopy.err
ls -la gpurun_out | grep r2af
