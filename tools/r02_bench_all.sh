#!/bin/bash
# round-2 bench lines of every BASELINE configuration (full-solve and fast-update dense moves) -> gpurun_out/r02_bench_*.json
set -u
run() { name=$1; shift; python bench.py "$@" > gpurun_out/r02_bench_$name.json 2> gpurun_out/r02_bench_$name.err || tail -3 gpurun_out/r02_bench_$name.err; }
run c5
run c5_ipr --measure-ipr --steps 2
for w in c1 c2 c3 c4t c4h; do
  run $w --workload $w
  run ${w}_fast --workload $w --fast-update --no-cpu-baseline
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02_bench_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); continue
    cb = d.get("cpu_baseline") or {}
    print("%-22s value %10.0f  e2e %10.0f  ms/step %8.2f  acc %.3f  dominant %-10s roofline %.3f (%s)  cpu port %s lapack %s" % (
        f.split("r02_bench_")[1][:-5], d["value"], (d.get("e2e") or {}).get("value", 0) or 0, d["ms_per_step"], d.get("accept_rate", 0), d["dominant_kernel"],
        d["roofline"]["frac"], d["roofline"]["kernel"], round(cb.get("value", 0)) if cb else None, round((cb.get("lapack") or {}).get("value", 0)) if cb else None))
PY
