#!/bin/bash
# usage: scripts/bench_all.sh <tag> [workloads...]  -- runs bench.py per workload on the GPU box, prints a one-line summary each
tag=$1; shift
wl=${@:-c2 c5 c5solved c4}
for w in $wl; do
  timeout 300 python bench.py --workload $w --steps 100 --warmup 5 --no-cpu > gpurun_out/${tag}_bench_$w.json 2> gpurun_out/${tag}_bench_$w.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_bench_$w.json"))
    r = d["roofline"]
    print("$w", "value %.4g e2e %.4g ms %.4f frac %.3f ach %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["frac"], r["achieved"]), {k: round(v, 4) for k, v in r["phase_ms"].items()}, d["gpu_launches"])
except Exception as e:
    print("$w FAILED", e); print(open("gpurun_out/${tag}_bench_$w.err").read()[-1500:])
PY
done
