#!/usr/bin/env python
"""One group of 100K-vertex spheres through hb_encode_batch on page-locked host buffers: wall time against the bytes that
cross the link, with the gathered upload (one kernel per upload stage reads the host arrays directly) and with one
cudaMemcpyAsync per array (HARRY_B200_GATHER_UPLOADS=0)."""
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
ap.add_argument("--meshes", type=int, default=195)
ap.add_argument("--distinct", type=int, default=4)
ap.add_argument("--reps", type=int, default=4)
args = ap.parse_args()
d = tempfile.mkdtemp()
loads = [bench.BatchMesh(*bench.BATCH_SHAPE, 100 + k, d) for k in range(args.distinct)]
ctx = capi.Context(0)
pinned = [bench.pin_mesh(b.raw) for b in loads]
meshes = [pinned[k % args.distinct] for k in range(args.meshes)]
req = [(1, loads[0].new_quant, loads[0].groups)]
os.environ["HARRY_B200_GROUP_HALF_EDGES"] = str(1 << 30)   # one group: upload, then kernels, then download
for mode in ("1", "0", "1", "0"):
    os.environ["HARRY_B200_GATHER_UPLOADS"] = mode
    best = 1e9
    for rep in range(args.reps):
        for m, b in zip(pinned, loads):
            m.lists[1].rows[...] = b.raw.lists[1].rows
            m.lists[1].quants = [0] * m.lists[1].ncomp
        up0 = ctx.h2d_bytes()
        t0 = time.perf_counter()
        streams, bounds, release = ctx.encode_batch(meshes, req, copy=False)
        dt = time.perf_counter() - t0
        up = ctx.h2d_bytes() - up0
        release()
        best = min(best, dt)
    print(f"gather {mode}: {args.meshes} meshes, {up / 1e9:.3f} GB up, best of {args.reps}: {best * 1e3:.2f} ms wall = {up / best / 1e9:.1f} GB/s including kernels and download")
