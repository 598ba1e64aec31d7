#!/bin/bash
# Round 2, GPU call AA (1 x B200): cudaLimitMaxL2FetchGranularity 32 / 64 / 128 (K4 moves 28.5 GB of DRAM for 7.1 GB algorithmic).
set -x
mkdir -p gpurun_out
python - <<'PY' > gpurun_out/r2aa_limit.txt 2>&1
import torch, ctypes
torch.cuda.init(); torch.zeros(1, device="cuda")
rt = ctypes.CDLL("libcudart.so.12")
v = ctypes.c_size_t(0); print("rc", rt.cudaDeviceGetLimit(ctypes.byref(v), 0x05), "default L2 fetch granularity", v.value)
PY
for v in 32 64 128; do
  MM_L2_FETCH=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2aa_fetch$v.json 2>> gpurun_out/r2aa.err
done
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2aa_default.json 2>> gpurun_out/r2aa.err
ls gpurun_out | grep r2aa
