#!/bin/bash
# Round 2, GPU call H (1 x B200): the final tree -- smoke, the full bench line (all N = 1 legs), the reference arm, the launch
# list and the ncu --set full capture the committed summaries come from (raw pages exported on the box; reports kept only if
# the whole output stays below gpurun's 64 MiB).
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h4_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h4_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h4_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2h4_smoke.log)
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h4_bench.json 2> gpurun_out/r2h4_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h4_bench_ref.json 2> gpurun_out/r2h4_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h4_launches.csv \
  python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2h4_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem|l1_filter_gather16|l1_probe_tma|sketch_blockmin_kernel' \
  -o /tmp/r2h4_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2h4_ncu.log 2>&1
ncu -i /tmp/r2h4_full.ncu-rep --page raw --csv > gpurun_out/r2h4_full_raw.csv 2>/dev/null
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'em_round' -s 60 -c 2 \
  -o /tmp/r2h4_em -f python bench.py --workload config4 --reads 40000 > gpurun_out/r2h4_em.log 2>&1
ncu -i /tmp/r2h4_em.ncu-rep --page raw --csv > gpurun_out/r2h4_em_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out | grep r2h3
