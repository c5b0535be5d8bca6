#!/bin/bash
# bench.py over the BASELINE.json configs on one GPU (run under gpurun); one JSON line per config into gpurun_out/
TAG=${1:-r2}
for c in c3 c3-resets c2 c4 c4-heavy c5; do
  python bench.py --config $c --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_${c}.json 2> gpurun_out/${TAG}_bench_${c}.err || echo "config $c failed"
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${c}.json"))
print("$c", d["config"]["workload"][:60], "| value %.3e ms %.3f e2e %.3e | episodes %s mean_len %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["episode_stats"]["episodes"], d["episode_stats"]["mean_length"]))
print("   roofline issue frac", d["roofline"]["frac"], "hbm step frac", d["roofline"]["hbm"]["step_frac"], "fp32 frac", d["roofline"]["fp32"]["frac"], "alg", d["roofline"]["fp32"]["algorithmic_frac"])
PY
done
