#!/usr/bin/env python
"""Per-call wall times of the host-buffer C ABI (the e2e path of bench.py): bounds, requant,
attr_encode, attr_decode, dequantize.  Usage under gpurun: python tools/e2e_probe.py [--nr --ns]"""
import argparse
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from harry_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=2237)
ap.add_argument("--ns", type=int, default=4472)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
w = bench.Workload(args.nr, args.ns, tempfile.mkdtemp(prefix="harry_e2e_"))
ctx = capi.Context(0)
vl = 1
hraw = bench.pin_mesh(w.raw)
hdec = bench.pin_mesh(w.dec)
praw = [la.rows.copy() for la in w.raw.lists]
pdec = [la.rows.copy() for la in w.dec.lists]
for it in range(args.reps):
    for la, src in zip(hraw.lists, praw):
        la.rows[...] = src
        la.quants = [0] * la.ncomp
    for la, src, ref in zip(hdec.lists, pdec, w.dec.lists):
        la.rows[...] = src
        la.quants = list(ref.quants)
    T = [time.perf_counter()]
    la = hraw.lists[vl]
    mn, mx = ctx.bounds(la); T.append(time.perf_counter())
    sc = bench.float_scale_row(la, mn, mx); T.append(time.perf_counter())
    ctx.requant(la, w.new_quant[vl], mn, sc); T.append(time.perf_counter())
    km0 = ctx.last_timing() if hasattr(ctx, "last_timing") else None
    streams, release = ctx.attr_encode_view(hraw); T.append(time.perf_counter())
    km1 = ctx.last_timing() if hasattr(ctx, "last_timing") else None
    ctx.attr_decode(hdec); T.append(time.perf_counter())
    km2 = ctx.last_timing() if hasattr(ctx, "last_timing") else None
    ld = hdec.lists[vl]
    ctx.requant(ld, [0] * ld.ncomp, w.dec_bounds[vl][0], w.dec_bounds[vl][2]); T.append(time.perf_counter())
    names = ["bounds", "scale(host)", "requant", "attr_encode", "attr_decode", "dequant"]
    print("iter", it, " ".join(f"{n}={1e3 * (b - a):.1f}ms" for n, a, b in zip(names, T[:-1], T[1:])), f"total={1e3 * (T[-1] - T[0]):.1f}ms",
          "enc(kernel,copy)=", km1, "dec(kernel,copy)=", km2, flush=True)
    del streams
    release()
