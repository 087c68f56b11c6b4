#!/usr/bin/env python
"""Device-resident / e2e lines for BASELINE configs[0], [2], [3] (bench.run_other_configs) without the rest of the bench."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
from harry_b200 import capi  # noqa: E402

ctx = capi.Context(0)
for line in bench.run_other_configs(ctx, tempfile.mkdtemp(), bench.peaks()[0]):
    print(json.dumps(line))
ctx.close()
