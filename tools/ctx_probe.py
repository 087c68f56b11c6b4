#!/usr/bin/env python
"""Where the fixed cost of a drop-in process goes: library load, bare CUDA context, hb_ctx_create, first calls, destroy."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
t = time.perf_counter()
lib = C.CDLL(os.path.join(ROOT, "harry_b200", "libharry_b200.so"))
print(f"dlopen libharry_b200.so      {time.perf_counter() - t:7.3f} s")
t = time.perf_counter()
rt = C.CDLL("libcudart.so.12")
rt.cudaFree(0)
print(f"cudaFree(0) (context)        {time.perf_counter() - t:7.3f} s")
h = C.c_void_p()
lib.hb_ctx_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
t = time.perf_counter()
rc = lib.hb_ctx_create(0, C.byref(h))
print(f"hb_ctx_create                {time.perf_counter() - t:7.3f} s  rc {rc}")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
from harry_b200 import capi, flatten, meshgen  # noqa: E402
ctx = capi.Context(0)
mesh = flatten.mesh_arrays(meshgen.uv_sphere(20, 40))
for rep in range(3):
    t = time.perf_counter()
    s = ctx.attr_encode(mesh)
    print(f"hb_attr_encode (853 vertices) {time.perf_counter() - t:7.3f} s")
t = time.perf_counter()
ctx.close()
lib.hb_ctx_destroy.argtypes = [C.c_void_p]
lib.hb_ctx_destroy(h)
print(f"hb_ctx_destroy x2            {time.perf_counter() - t:7.3f} s")
t = time.perf_counter()
rt.cudaDeviceReset()
print(f"cudaDeviceReset              {time.perf_counter() - t:7.3f} s")
