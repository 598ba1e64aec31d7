#!/bin/bash
# Round 2, GPU call AH (1 x B200): 32-byte slots with the first 8 contig ids of a hash's list inline (K4 pass 0 reads them from the slot's own sector).
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2ah_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ah_tests.log)
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2ah_bench.json 2> gpurun_out/r2ah_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off \
    -k regex:'l1_probe_filter' --csv --log-file gpurun_out/r2ah_ncu.csv python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > /dev/null 2>&1
ls gpurun_out | grep r2ah
