#!/usr/bin/env python
"""Turns ncu outputs brought back in gpurun_out/ into the small text summaries committed under
profiles/ (launch list -> per-kernel share of a step; --set full capture -> key metrics)."""
import csv
import subprocess
import sys
from collections import OrderedDict


def launches(path, out):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    h = rows[hdr]
    ci = {n: i for i, n in enumerate(h)}
    agg = OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) <= ci["Metric Value"] or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit in ("us", "usecond") else v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    tot = sum(v[1] for v in agg.values())
    with open(out, "w") as f:
        f.write("# kernel, launches, total_ms, share  (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{n},{ms:.4f},{ms / tot:.4f}\n")
    print("wrote", out, "total ms", round(tot, 2))


def full(rep, out, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h = rows[0]
    with open(out, "w") as f:
        for r in rows[2:]:
            f.write("## " + r[h.index("Kernel Name")] + "\n")
            for k in keys:
                if k in h:
                    f.write(f"{k} = {r[h.index(k)]} {rows[1][h.index(k)]}\n")
    print("wrote", out)


KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]

if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], KEYS)
