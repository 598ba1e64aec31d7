#!/bin/bash
# Round 2, GPU call C (2 x B200): NCCL paths (contig-shard exchange on the device, EM all-reduce, streamed chunks), N = 2 bench line
# with the extra legs; on one GPU: grouped L1 filter vs the flat one, host CLI phase timing.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/r2c_gpu.txt
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or map_ or fallback or variants or staged or cli_matches_golden" > gpurun_out/r2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests.log)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus 2 --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2c_config5_n2.json 2> gpurun_out/r2c_config5_n2.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
MM_L1_FILTER=flat timeout 300 python bench.py --steps 6 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2c_bench_flat.json 2>> gpurun_out/r2c_var.err
ls -la gpurun_out | grep r2c
