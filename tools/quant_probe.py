#!/usr/bin/env python
"""Times the quantization kernels (bounds + scale, requant, dequant) on a 10M x 3 float list, device resident."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from harry_b200 import capi, flatten, meshgen  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 9999394
pm = meshgen.uv_sphere(6, 7)
mesh = flatten.mesh_arrays(pm)
rng = np.random.default_rng(1)
rows = rng.standard_normal((n, 3)).astype(np.float32)
mesh.lists[1] = capi.make_list([rows[:, 0], rows[:, 1], rows[:, 2]], [capi.FLOAT] * 3, capi.T_VTX, [0, 0, 0])
mesh.bind_vtx = np.arange(mesh.nv, dtype=np.uint32)
ctx = capi.Context(0)
dm = capi.DeviceMesh(ctx, mesh)
dm.snapshot()
for rep in range(4):
    dm.restore()
    ctx.sync()
    ctx.profile(True)
    dm.quantize(1, [14, 14, 14], [0, 0, 0])
    dm.dequantize(1)
    ctx.sync()
    prof = ctx.profile_report()
    ctx.profile(False)
A = 3 * n
print(os.environ.get("HARRY_B200_FLAT_BLOCKS_PER_SM", "8"), {k: (round(ms * 1e3, 1), "us", round(A * (4 if "bounds" in k else 8) / (ms * 1e-3) / 1e9), "GB/s actual") for k, (c, ms) in prof.items()})
