#!/usr/bin/env python
"""bench.py lines of every BASELINE.json config (written by profiles/run_configs_1gpu.sh) -> markdown table + the raw lines.
  python profiles/make_config_table.py gpurun_out r2c profiles/r2_bench_configs.md"""
import json
import os
import sys

src, tag, out = sys.argv[1], sys.argv[2], sys.argv[3]
rows, raw = [], []
for c in ("c2", "c3", "c3-resets", "c4", "c4-heavy", "c5"):
    p = os.path.join(src, f"{tag}_bench_{c}.json")
    if not os.path.exists(p):
        continue
    d = json.loads([l for l in open(p) if l.startswith("{")][-1])
    r, f = d["roofline"], d["roofline"]["fp32"]
    es = d["episode_stats"]
    rows.append(f"| `{c}` | {d['config']['what']} | {d['value']:.3e} | {d['ms_per_step']:.3f} | {d['e2e']['value']:.3e} | {d['e2e']['ms_per_step']:.3f} | "
                f"{('%.3f' % r['frac']) if r['frac'] else '—'} | {r['hbm']['step_frac']:.4f} | {f['frac']:.4f} / {f['algorithmic_frac']:.4f} | "
                f"{int(es['episodes'])} ({es['mean_length']:.0f}) | {int(es['overflow'])} |")
    raw.append(json.dumps(d))
with open(out, "w") as fh:
    fh.write("# `bench.py --config` on one B200: every BASELINE.json config (round 2)\n\n"
             "`bash profiles/run_configs_1gpu.sh <tag>` under gpurun (20 timed steps after 5 warm-up and 100 settle steps, not under a profiler); "
             "value = device-resident env-steps/s, e2e = through `mrp_step_host` with pinned host buffers.  The issue fraction is only "
             "reported for the workload the committed ncu capture describes (Heavy-v0).\n\n"
             "| config | what | value (env-steps/s) | ms / step | e2e (env-steps/s) | e2e ms / step | issue frac | HBM algorithmic frac | FP32 executed / algorithmic frac | episodes finished in the timed region (mean length) | contact overflows |\n"
             "|---|---|---|---|---|---|---|---|---|---|---|\n" + "\n".join(rows) + "\n\nRaw lines:\n\n```\n" + "\n".join(raw) + "\n```\n")
print("\n".join(rows))
