#!/bin/bash
# Round 2, GPU call AC (2 x B200): final tree -- NCCL shard tests, N = 2 bench line with all extra legs, config-5 slice streamed through 2 GPUs.
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or map_golden or map_config1 or staged or ragged" > gpurun_out/r2ac_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ac_tests.log)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/r2ac_bench_n2.json 2> gpurun_out/r2ac_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus 2 --workload config5-slice --steps 2 --warmup 1 > gpurun_out/r2ac_config5_n2.json 2> gpurun_out/r2ac_config5_n2.err
ls -la gpurun_out | grep r2ac
