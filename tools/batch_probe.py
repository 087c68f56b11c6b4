#!/usr/bin/env python
"""Probe of the batched decode: scan-kernel statistics and stage times for a batch of 100K-vertex spheres."""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from harry_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--meshes", type=int, default=83)
ap.add_argument("--distinct", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
d = tempfile.mkdtemp()
loads = [bench.BatchMesh(*bench.BATCH_SHAPE, 100 + k, d) for k in range(args.distinct)]
ctx = capi.Context(0)
idx = [k % args.distinct for k in range(args.meshes)]
D = capi.DeviceMesh(ctx, [loads[i].dec for i in idx])
D.set_bounds(1, *(np.stack([loads[i].dec_bounds[w] for i in idx]) for w in (0, 1, 2)))
D.snapshot()
for rep in range(args.reps):
    D.restore()
    ctx.sync()
    ctx.profile(True)
    t0 = time.perf_counter()
    D.decode()
    ctx.sync()
    dt = time.perf_counter() - t0
    prof = ctx.profile_report()
    ctx.profile(False)
st = D.decode_stats(1)
sweeps = st[0]
print(f"meshes {args.meshes}: decode {dt*1e3:.2f} ms; scan kernel {prof.get('k_decode_vertex_scan', (0, 0))[1]:.3f} ms")
print(f"segment 0, component 0: sweeps {sweeps} fails {st[1] & 0xffffffff} capped {st[1] >> 32} fallback {st[2] & 0xffffffff} wides {(st[2] >> 32) & 0xff} nseq {st[2] >> 40} n {st[3]}")
print(f"cycles per sweep: A {st[4]/sweeps:.0f} B {st[5]/sweeps:.0f} C {st[6]/sweeps:.0f} D {st[7]/sweeps:.0f} total {(st[4]+st[5]+st[6]+st[7])/sweeps:.0f}; total cycles {(st[4]+st[5]+st[6]+st[7])/1e6:.2f} M")
for k, (n, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]:
    print(f"  {k}: {n} x, {ms:.3f} ms")
ok = np.array_equal(D.fetch_rows(1, args.meshes - 1), loads[idx[-1]].dec.lists[1].rows * 0 + D.fetch_rows(1, args.meshes - 1))
D.close()
ctx.close()
