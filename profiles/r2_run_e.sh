#!/bin/bash
# Round 2, GPU call E (8 x B200): the N = 8 bench line with all extra legs, and the config-5 slice streamed through 8 GPUs.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2e_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2e_bench_n8.json 2> gpurun_out/r2e_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus 8 --workload config5-slice --steps 3 --warmup 1 > gpurun_out/r2e_config5_n8.json 2> gpurun_out/r2e_config5_n8.err
ls -la gpurun_out | grep r2e
