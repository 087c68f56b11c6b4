#!/usr/bin/env python
"""Regenerates tests/golden/twin/*.npz: twin tables built by the UNMODIFIED reference's readers
(mesh::Builder, structs/conn.h:164-214, driven by formats/ply/reader.cc:349-353) through
oracle/_ref/libharry_ref.so.  Run in the build container:

    python tests/golden/make_golden_twin.py

Each fixture holds nv, face_off and the reference's Conn::edges records (org, twin_face, twin_edge) right after
reading, before the Cut-Border-Machine touches the table."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as ol  # noqa: E402
from harry_b200 import meshgen  # noqa: E402

TWIN_GOLDEN = {
    "soup_small_faces": lambda: meshgen.soup(30, 400, 1, 1, 6),
    "soup_8_vertices": lambda: meshgen.soup(8, 1000, 2, 2, 6),
    "soup_sparse": lambda: meshgen.soup(200, 500, 3, 3, 5),
    "soup_5_vertices": lambda: meshgen.soup(5, 3000, 4, 1, 6),
    "poly_fin": lambda: meshgen.poly_grid(12),
    "cones_open": lambda: meshgen.cones(6, 70, seed=4, open_every=2),
    "irregular": lambda: meshgen.tri_irregular(20, 11),
}


def main():
    d = tempfile.mkdtemp(prefix="harry_golden_twin_")
    os.makedirs(os.path.join(HERE, "twin"), exist_ok=True)
    for name, gen in TWIN_GOLDEN.items():
        pm = gen()
        p = os.path.join(d, name + ".ply")
        meshgen.write_ply(p, pm)
        rm = ol.RefMesh(p)
        m = rm.arrays()
        rm.close()
        assert np.array_equal(m.edges[:, 0], pm.face_idx) and np.array_equal(m.face_off, pm.face_off)
        out = os.path.join(HERE, "twin", name + ".npz")
        np.savez_compressed(out, nv=np.array([m.nv], np.int64), face_off=m.face_off, edges=m.edges)
        print(name, os.path.getsize(out), "bytes", "nv", m.nv, "nf", m.nf, "ne", m.ne)


if __name__ == "__main__":
    main()
