#!/bin/bash
# Round 2, GPU call AG (1 x B200): DRAM bytes of K4 / K5a / K5b under cudaLimitMaxL2FetchGranularity 32 against the default (64).
set -x
mkdir -p gpurun_out
for v in 32 64; do
  MM_L2_FETCH=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none --profile-from-start off \
    -k regex:'l1_probe_filter|l2_classify_smem|l2_sweep_band|sketch_blockmin' --csv --log-file gpurun_out/r2ag_ncu$v.csv python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > /dev/null 2>&1
done
ls gpurun_out | grep r2ag
