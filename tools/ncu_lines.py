#!/usr/bin/env python
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` export per CUDA source line:
stall samples, executed warp instructions, and the dominant stall reasons.  usage: ncu_lines.py file.csv [top]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = defaultdict(lambda: {"samples": 0, "inst": 0, "stalls": defaultdict(int), "src": ""})
fname, hdr, cur = None, None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
        stall_cols = [(k, h) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0].strip():
        cur = (fname, int(r[0]))
        agg[cur]["src"] = r[1].strip()[:110]
    if cur is None or not r[2].strip():
        continue
    a = agg[cur]
    try:
        a["samples"] += int(r[si] or 0)
        a["inst"] += int(r[ii] or 0)
        for k, h in stall_cols:
            v = int(r[k] or 0)
            if v:
                a["stalls"][h[6:]] += v
    except ValueError:
        pass
tot_s = sum(a["samples"] for a in agg.values())
tot_i = sum(a["inst"] for a in agg.values())
print(f"total samples {tot_s}, warp instructions {tot_i}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
    st = ", ".join(f"{k} {v}" for k, v in sorted(a["stalls"].items(), key=lambda kv: -kv[1])[:3])
    print(f"{key[0]}:{key[1]:5d}  samples {100*a['samples']/tot_s:5.1f}%  inst {100*a['inst']/tot_i:5.1f}%  [{st}]  {a['src']}")
