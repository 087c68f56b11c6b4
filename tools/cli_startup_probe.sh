#!/bin/bash
# fixed cost of one drop-in CLI process (CUDA context creation, library load) against the reference CLI on a tiny mesh
cd "$(dirname "$0")/.."
nvidia-smi -q | grep -i -m1 "persistence mode"
python - <<'PY'
import sys, os, time, subprocess, tempfile
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from harry_b200 import meshgen
d = tempfile.mkdtemp()
ply = os.path.join(d, "s.ply")
meshgen.write_ply(ply, meshgen.uv_sphere(20, 40))
for name, b in (("reference", "oracle/_ref/harry"), ("harry_b200", "harry_b200/host/bin/harry_b200")):
    for rep in range(3):
        t0 = time.perf_counter()
        r = subprocess.run([b, ply, os.path.join(d, name + ".hry"), "-l1", "-q14"], capture_output=True, text=True)
        print(name, "rep", rep, "rc", r.returncode, f"{time.perf_counter() - t0:.3f} s")
t0 = time.perf_counter()
import ctypes
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaFree(0)
print(f"cudaFree(0) in this process: {time.perf_counter() - t0:.3f} s")
PY
