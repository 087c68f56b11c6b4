#!/usr/bin/env python
"""Small driver for ncu captures: prepares the workload, then runs `--reps` device-resident
encode+decode steps.  Usage under gpurun:
    ncu --set full --clock-control none --import-source on -k regex:k_encode_main -s 1 -c 1 \
        -o gpurun_out/prof python tools/profile_driver.py --nr 708 --ns 1412
"""
import argparse
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import bench  # noqa: E402
from harry_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--nr", type=int, default=708)
ap.add_argument("--ns", type=int, default=1412)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-decode", action="store_true")
ap.add_argument("--verify", action="store_true", help="compare the encode streams and the decoded rows with the reference's at this size")
args = ap.parse_args()
w = bench.Workload(args.nr, args.ns, tempfile.mkdtemp(prefix="harry_prof_"), keep_expected=args.verify)
ctx = capi.Context(0)
E = capi.DeviceMesh(ctx, w.raw)
E.snapshot()
D = capi.DeviceMesh(ctx, w.dec)
for l, (mn, mx, sc) in enumerate(w.dec_bounds):
    if w.dec.lists[l].ncomp:
        D.set_bounds(l, mn, mx, sc)
D.snapshot()
for _ in range(args.reps):
    E.restore()
    D.restore()
    E.quantize(1, w.new_quant[1], w.raw.lists[1].groups)
    E.encode()
    if args.verify:
        ok, why = E.fetch_streams().equal(w.enc_expected)
        print("verify: encode streams (types, history offsets, residual symbols, histograms) == reference:", ok, why, flush=True)
    if not args.no_decode:
        D.decode()
        if args.verify:
            import numpy as np
            good = all(np.array_equal(D.fetch_rows(l), exp) for l, exp in enumerate(w.dec_expected) if w.dec.lists[l].ncomp)
            print("verify: decoded rows == reference decoder output:", good, flush=True)
        D.dequantize(1)
ctx.sync()
print("done", ctx.launches(), "decode stats [sweeps, hyp, plain, hyp_adv, cyc_load, cyc_traj, cyc_resolve, cyc_write+plain]:", D.decode_stats(1)[:8])
