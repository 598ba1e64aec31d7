#!/bin/bash
# Round 2, GPU call AB (1 x B200): K4's random loads as ld.global.nc / ld.global.cg / ld.global.nc.L1::no_allocate (L2 sees 3.4 sectors per sector the loads ask for),
# with the DRAM bytes of the kernel from ncu.
set -x
mkdir -p gpurun_out
for v in 0 1 2; do
  MM_K4_LD=$v timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2ab_ld$v.json 2>> gpurun_out/r2ab.err
  MM_K4_LD=$v timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum --clock-control none --profile-from-start off \
    -k regex:'l1_probe_filter' --csv --log-file gpurun_out/r2ab_ncu$v.csv python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > /dev/null 2>&1
done
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "map_golden or config1 or at_scale" > gpurun_out/r2ab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2ab_tests.log)
ls gpurun_out | grep r2ab
