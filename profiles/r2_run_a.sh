#!/bin/bash
# Round 2, GPU call A (1 x B200): parity tests, smoke, the bench line, K5b occupancy variants, launch list + ncu --set full.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.txt
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2a_smoke.log)
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
for v in "256 4 0" "128 8 0" "128 4 0" "128 4 1"; do
  set -- $v
  if [ "$3" = "1" ]; then export MM_SWEEP_2CTA=1; else unset MM_SWEEP_2CTA; fi
  MM_SWEEP_BAND=$1 MM_SWEEP_RING=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2a_bench_b$1_r$2_c$3.json 2>> gpurun_out/r2a_var.err
done
unset MM_SWEEP_2CTA
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2a_launches.csv \
  python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2a_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem|l1_filter_gather16|l1_probe_tma|em_round|sketch_blockmin_kernel|read_sketch_block_kernel|l2_strand' \
  -o gpurun_out/r2a_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2a_ncu.log 2>&1
timeout 600 python bench.py --workload config4-small > gpurun_out/r2a_config4small.json 2> gpurun_out/r2a_config4small.err
ls -la gpurun_out | tail -30
