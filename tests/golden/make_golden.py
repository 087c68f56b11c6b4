#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libharry_ref.so, built
from /root/reference by oracle/Makefile).  Run in the build container:

    python tests/golden/make_golden.py

Every array in a fixture is an output of the reference itself: its readers, quant::set_bounds /
set_scale / requant, cbm::encode (traversal order, final twin table), AttrCoder<Capture> (symbol
streams), the real models' final frequency tables, the .hry bytes, and on the decode side what the
real AttrDecoder read from that file and reconstructed."""
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import cases  # noqa: E402
import golden_io  # noqa: E402
from harry_b200 import meshgen  # noqa: E402

GOLDEN = {
    "sphere_q14": (lambda d: cases._ply(d, "g_s.ply", meshgen.uv_sphere(12, 20, noise_seed=2)), [(1, -1, 14)]),
    "sphere_lossless": (lambda d: cases._ply(d, "g_s.ply", meshgen.uv_sphere(12, 20, noise_seed=2)), []),
    "poly_q10": (lambda d: cases._ply(d, "g_p.ply", meshgen.poly_grid(7)), [(1, -1, 10)]),
    "poly_lossless": (lambda d: cases._ply(d, "g_p.ply", meshgen.poly_grid(7)), []),
    "obj_multi": (lambda d: _obj(d), [(0, -1, 12), (1, -1, 9), (2, -1, 10), (3, -1, 11)]),
    # integer / mixed source types (round 2): uchar colours requantized next to quantized coordinates; int / short / uchar
    # properties at misaligned offsets on an irregular triangulation, quantized from their integer types, and lossless
    "rgb_q5_xyz_q14": (lambda d: cases._ply(d, "g_rgb.ply", meshgen.with_typed_props(meshgen.uv_sphere(11, 17, noise_seed=8), seed=3,
                       vtx=(("red", "uint8"), ("green", "uint8"), ("blue", "uint8")))), [(1, 0, 14), (1, 1, 14), (1, 2, 14), (1, 3, 5), (1, 4, 5), (1, 5, 5)]),
    "ints_irr_q": (lambda d: cases._ply(d, "g_int.ply", _ints(d)), [(1, 0, 10), (1, 1, 10), (1, 2, 10), (1, 3, 5), (1, 4, 13), (1, 5, 7), (0, 0, 9)]),
    "ints_irr_lossless": (lambda d: cases._ply(d, "g_int.ply", _ints(d)), []),
}


def _ints(d):
    return meshgen.with_typed_props(meshgen.tri_irregular(14, 9), seed=6, vtx=(("material", "int32"), ("temp", "int16"), ("red", "uint8")), face=(("group", "int16"),))


def _obj(d):
    p = os.path.join(d, "g_o.obj")
    if not os.path.exists(p):
        meshgen.write_obj_latlong(p, 10, 9, multi_region=True)
    return p


def main():
    d = tempfile.mkdtemp(prefix="harry_golden_")
    for name, (gen, loq) in GOLDEN.items():
        c = cases.Case(d, "golden_" + name, gen, loq)
        out = os.path.join(HERE, name + ".npz")
        golden_io.save_case(c, out)
        print(name, os.path.getsize(out), "bytes", "nv", c.raw.nv, "hry", os.path.getsize(c.hry_path))


if __name__ == "__main__":
    main()
