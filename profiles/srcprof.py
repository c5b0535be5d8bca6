#!/usr/bin/env python
"""Rank the CUDA source lines of one kernel by warp-stall samples, from an `ncu --set full --import-source on` report
(read here, no GPU needed):
  ncu -i REP --page source --print-source cuda,sass --csv --kernel-name regex:k_pre > /tmp/k_pre.csv
  python profiles/srcprof.py /tmp/k_pre.csv [top_n]
"""
import csv
import sys

path = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, head, recs = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        head = r[:]
        head[1] = "Src"          # two columns are called Source: the CUDA line and the SASS text
        continue
    if head is None or r[0] == "":
        continue                 # SASS rows (and "..." separators) belong to the source row above them
    d = dict(zip(head, r))
    d["file"] = cur
    recs.append(d)


def num(x):
    try:
        return float(x)
    except (TypeError, ValueError):
        return 0.0


tot = sum(num(d.get("# Samples")) for d in recs) or 1.0
toti = sum(num(d.get("Instructions Executed")) for d in recs) or 1.0
print("total samples", tot, "total warp instructions", toti)
byfile = {}
for d in recs:
    byfile[d["file"]] = byfile.get(d["file"], 0) + num(d.get("# Samples"))
print({k: round(v / tot, 3) for k, v in byfile.items()})
recs.sort(key=lambda d: -num(d.get("# Samples")))
for d in recs[:topn]:
    s = num(d["# Samples"])
    print(f"{d['file']}:{d['Line No']:>5} smp {s / tot * 100:5.2f}% inst {num(d['Instructions Executed']) / toti * 100:5.2f}% "
          f"thr/inst {num(d.get('Avg. Threads Executed')):4.1f} long_sb {num(d.get('stall_long_sb')) / max(s, 1):.2f} | {d['Src'].strip()[:100]}")
