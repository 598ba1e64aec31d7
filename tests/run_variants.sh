#!/bin/bash
# usage: tests/run_variants.sh "<label> <ENV=..> ..." ...   -- one short bench per variant, one summary line each
for spec in "$@"; do
  set -- $spec; label=$1; shift
  env "$@" timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$label.log 2>&1
  python - "$label" <<'PY'
import json, sys
lab = sys.argv[1]
try:
    l = [x for x in open(f"gpurun_out/bench_{lab}.log") if x.startswith("{")][-1]; d = json.loads(l)
    st = {k.replace("_ms", ""): round(v, 2) for k, v in d["roofline"]["stage_ms"].items()}
    print(lab, "ms/step %.1f value %.0f e2e %.0f kernel_ms %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["roofline"]["kernel_ms_per_step"]), st, "fast", d["config"]["smem_swept"], "/", d["config"]["candidates_per_step"])
except Exception as e:
    print(lab, "FAILED", e); print(open(f"gpurun_out/bench_{lab}.log").read()[-800:])
PY
done
