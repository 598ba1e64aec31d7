#!/bin/bash
# Round 2, GPU call P (1 x B200): pruned K5b, window starts per segment (the longest single item is the kernel's critical path).
set -x
mkdir -p gpurun_out
for v in 256 512 1024 2048; do
  MM_SWEEP_SEG=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2p_seg$v.json 2>> gpurun_out/r2p.err
done
ls gpurun_out | grep r2p
