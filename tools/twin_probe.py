#!/usr/bin/env python
"""hb_twin_match at the BASELINE configs[1] size: kernel / copy / wall times per call and the per-kernel split
(target of the ncu captures under profiles/).  --permute renumbers the vertices at random (no locality between
the file order of the faces and the vertex indices)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from harry_b200 import capi, meshgen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nr", type=int, default=2237)
    ap.add_argument("--ns", type=int, default=4472)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--permute", action="store_true")
    args = ap.parse_args()
    pm = meshgen.uv_sphere(args.nr, args.ns)
    org = pm.face_idx
    if args.permute:
        org = np.random.default_rng(1).permutation(pm.nv).astype(np.uint32)[org]
    ctx = capi.Context(0)
    ne = org.shape[0]
    out = ctx.twin_match(pm.nv, pm.face_off, org)           # warm-up; host buffers are pageable here (bench.py pins them)
    ctx.profile(True)
    kern = wall = 0.0
    for _ in range(args.reps):
        t0 = time.perf_counter()
        ctx.twin_match(pm.nv, pm.face_off, org, out=out)
        wall += time.perf_counter() - t0
        kern += ctx.timing()[0]
    prof = ctx.profile_report()
    ctx.profile(False)
    print(f"permute={args.permute} ne={ne} kernel_ms={kern / args.reps:.3f} wall_ms={wall / args.reps * 1e3:.1f} "
          + " ".join(f"{k}={ms / n:.3f}" for k, (n, ms) in prof.items()), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
