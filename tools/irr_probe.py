#!/usr/bin/env python
"""Vertex-decode statistics on an irregular triangulation (meshgen.tri_irregular): sweeps, failed
boundaries, fallbacks of the scan kernel, and parity of the decoded rows against the reference."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from harry_b200 import capi, meshgen  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=700)
ap.add_argument("--q", type=int, default=14)
args = ap.parse_args()
d = tempfile.mkdtemp(prefix="harry_irr_")
c = cases.Case(d, "irr", lambda dd: cases._ply(dd, "irr.ply", meshgen.tri_irregular(args.n, 3)), [(1, -1, args.q)])
ctx = capi.Context(0)
m = c.decode_input()
dm = capi.DeviceMesh(ctx, m)
for l, (mn, mx) in enumerate(c.dec_bounds):
    if m.lists[l].ncomp:
        dm.set_bounds(l, mn, mx, c.deq_scale[l])
dm.snapshot()
for rep in range(2):
    dm.restore()
    ctx.sync()
    t0 = time.perf_counter()
    dm.decode()
    ctx.sync()
    dt = time.perf_counter() - t0
ok = np.array_equal(dm.fetch_rows(1), c.dec.lists[1].rows)
st = dm.decode_stats(1)
print(f"irregular n={args.n} q={args.q}: {c.dec.nv} vertices, decode {dt * 1e3:.2f} ms, parity {ok}, sweeps {st[0]}, fails {st[1] & 0xffffffff}, capped {st[1] >> 32}, "
      f"fallbacks {st[2] & 0xffffffff}, wides {(st[2] >> 32) & 0xff}, seq stretches {st[2] >> 40}, cycles A/B/C/D {st[4:8]}")
