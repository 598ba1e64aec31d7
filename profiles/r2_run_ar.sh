#!/bin/bash
# Round 2, GPU call AR (2 x B200): the committed tree -- NCCL shard tests and the N = 2 bench line with its extra legs.
set -x
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_gpu_parity.py -m gpu -x -q -k "multi or map_golden or map_config1 or staged or ragged" > gpurun_out/r2ar_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ar_tests.log)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2ar_bench_n2.json 2> gpurun_out/r2ar_bench_n2.err
ls -la gpurun_out | grep r2ar
