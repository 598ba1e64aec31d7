#!/bin/bash
# Round 2, GPU call Z (1 x B200): the final tree -- all GPU tests, smoke, the full bench line (all N = 1 legs), the reference arm,
# the launch list and the ncu --set full captures (raw + source pages) the committed summaries come from.
set -x
mkdir -p gpurun_out
T=${1:-r2z}
(timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log)
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log)
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/${T}_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'l2_sweep_band|l2_classify_smem|l2_prune_warp|l1_probe_filter|sketch_blockmin_kernel|read_sketch_block_kernel' \
  -o /tmp/${T}_full -f python bench.py --profile-step --warmup 2 --no-extras --no-cpu-baseline > gpurun_out/${T}_ncu.log 2>&1
ncu -i /tmp/${T}_full.ncu-rep --page raw --csv > gpurun_out/${T}_full_raw.csv 2>/dev/null
ncu -i /tmp/${T}_full.ncu-rep --page source --csv > gpurun_out/${T}_source.csv 2>/dev/null
timeout 900 python bench.py --workload config4 --reads 100000 > gpurun_out/${T}_config4_100k.json 2> gpurun_out/${T}_config4_100k.err
du -sh gpurun_out; ls -la gpurun_out | grep ${T}
