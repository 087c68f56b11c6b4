#!/usr/bin/env python
"""BASELINE configs[2] at scale: OBJ lat-long sphere with v / vt / vn (corner lists, LHIST / HIST emissions),
-l0 -q14 -l2 -q10: GPU encode + decode times (device resident) and parity against the reference."""
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
ap.add_argument("--nr", type=int, default=400)
ap.add_argument("--ns", type=int, default=600)
ap.add_argument("--multi", action="store_true")
args = ap.parse_args()
d = tempfile.mkdtemp(prefix="harry_obj_")


def gen(dd):
    p = os.path.join(dd, "o.obj")
    meshgen.write_obj_latlong(p, args.nr, args.ns, multi_region=args.multi)
    return p


t0 = time.perf_counter()
c = cases.Case(d, "obj", gen, [(0, -1, 14), (2, -1, 10)])
print(f"reference side prepared in {time.perf_counter() - t0:.1f} s: {c.enc.nv} vertices, {c.enc.nf} faces, lists "
      f"{[(l.nrows, l.ncomp) for l in c.enc.lists]}", flush=True)
ctx = capi.Context(0)
# encode (device resident)
E = capi.DeviceMesh(ctx, c.enc)
for rep in range(2):
    ctx.sync()
    t0 = time.perf_counter()
    E.encode()
    ctx.sync()
    te = time.perf_counter() - t0
ok_e, why = E.fetch_streams().equal(c.enc_streams)
# decode
m = c.decode_input()
D = capi.DeviceMesh(ctx, m)
for l, (mn, mx) in enumerate(c.dec_bounds):
    if m.lists[l].ncomp:
        D.set_bounds(l, mn, mx, c.deq_scale[l])
D.snapshot()
for rep in range(2):
    D.restore()
    ctx.sync()
    t0 = time.perf_counter()
    D.decode()
    ctx.sync()
    td = time.perf_counter() - t0
ok_d = all(np.array_equal(D.fetch_rows(l), la.rows) for l, la in enumerate(c.dec.lists) if la.ncomp)
nattr = sum(l.nrows * l.ncomp for l in c.enc.lists)
print(f"OBJ {args.nr}x{args.ns}: encode {te * 1e3:.2f} ms (parity {ok_e} {why}), decode {td * 1e3:.2f} ms (parity {ok_d}), "
      f"{nattr} attrs -> {nattr / (te + td) / 1e6:.1f} M attr/s", flush=True)
