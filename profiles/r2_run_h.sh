#!/bin/bash
# Round 2, GPU call H (1 x B200): the final tree -- parity tests, smoke, the full bench line (all N = 1 legs), the reference arm,
# the launch list and the ncu --set full capture the committed summaries come from.
set -x
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2h_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2h_smoke.log)
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2h_bench_ref.json 2> gpurun_out/r2h_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2h_launches.csv \
  python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2h_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem|l1_filter_gather16|l1_probe_tma|sketch_blockmin_kernel|l2_strand|em_round|read_sketch_block' \
  -o gpurun_out/r2h_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/r2h_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'em_round' -s 60 -c 3 \
  -o gpurun_out/r2h_em -f python bench.py --workload config4 --reads 40000 > gpurun_out/r2h_em.log 2>&1
ls -la gpurun_out | grep r2h
