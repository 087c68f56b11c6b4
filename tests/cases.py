"""Shared test cases: the small BASELINE configs as flattened meshes (numpy), built through the
reference harness (oracle/_ref/libharry_ref.so: the unmodified reference's readers, quantizer and
Cut-Border-Machine traversal).  Test infrastructure only."""
from __future__ import annotations

import os

import numpy as np

import oracle_lib as ol
from harry_b200 import capi, meshgen

# name -> (generator, quantization triples (list, component or -1, bits))
CASES = {
    "sphere_lossless": (lambda d: _ply(d, "s.ply", meshgen.uv_sphere(40, 80)), []),
    "sphere_q14": (lambda d: _ply(d, "s.ply", meshgen.uv_sphere(40, 80)), [(1, -1, 14)]),
    "sphere_noise_q12": (lambda d: _ply(d, "sn.ply", meshgen.uv_sphere(33, 51, noise_seed=3)), [(1, -1, 12)]),
    "sphere_q8_mixed": (lambda d: _ply(d, "s2.ply", meshgen.uv_sphere(20, 31, noise_seed=5)), [(1, 0, 8), (1, 2, 20)]),
    "poly_lossless": (lambda d: _ply(d, "p.ply", meshgen.poly_grid(24)), []),
    "poly_q10": (lambda d: _ply(d, "p.ply", meshgen.poly_grid(24)), [(1, -1, 10), (0, -1, 9)]),
    "obj_q14_q10": (lambda d: _obj(d, "o.obj", False), [(0, -1, 14), (2, -1, 10)]),
    "obj_lossless": (lambda d: _obj(d, "o.obj", False), []),
    "obj_multi_q14": (lambda d: _obj(d, "om.obj", True), [(0, -1, 14)]),
    # the integer storage types of the vertex decoder's scan kernel (u8 / u16 / u32), regular and irregular meshes
    "sphere_q8": (lambda d: _ply(d, "sn8.ply", meshgen.uv_sphere(33, 51, noise_seed=4)), [(1, -1, 8)]),
    "sphere_q20": (lambda d: _ply(d, "sn20.ply", meshgen.uv_sphere(29, 47, noise_seed=6)), [(1, -1, 20)]),
    "irr_q14": (lambda d: _ply(d, "irr.ply", meshgen.tri_irregular(48, 11)), [(1, -1, 14)]),
    "irr_q7": (lambda d: _ply(d, "irr.ply", meshgen.tri_irregular(48, 11)), [(1, -1, 7)]),
    "irr_q24": (lambda d: _ply(d, "irr.ply", meshgen.tri_irregular(48, 11)), [(1, -1, 24)]),
    # 80 wide fans (more than the 64 the wide-fan path takes at once), quantized and lossless float
    "cones_q12": (lambda d: _ply(d, "cones.ply", meshgen.cones(80, 90)), [(1, -1, 12)]),
    "cones_lossless": (lambda d: _ply(d, "cones.ply", meshgen.cones(80, 90)), []),
    # open wide fans (apex on a border): forward part, then backward part of the fan walk
    "cones_open_q12": (lambda d: _ply(d, "coneso.ply", meshgen.cones(40, 100, seed=4, open_every=2)), [(1, -1, 12)]),
    "cones_open_lossless": (lambda d: _ply(d, "coneso.ply", meshgen.cones(40, 100, seed=4, open_every=2)), []),
    "irr_big_q12": (lambda d: _ply(d, "irrb.ply", meshgen.tri_irregular(160, 5)), [(1, -1, 12)]),
    "obj_multi_all": (lambda d: _obj(d, "om.obj", True), [(0, -1, 12), (1, -1, 9), (2, -1, 10), (3, -1, 11)]),
    # native integer / mixed source types (structs/mixing.h:18-19, quant.h:99-112,168, prediction.h:47-63,82-99; SURVEY
    # Appendix C.13): uchar colours next to float coordinates -- quantization selected per component (`-a`), the colours
    # lossless (mixed storage types in one list) or requantized from their integer type (rescale_int)
    "rgb_xyz_q12": (lambda d: _ply(d, "rgb.ply", _rgb_sphere()), [(1, 0, 12), (1, 1, 12), (1, 2, 12)]),
    "rgb_q5_xyz_q14": (lambda d: _ply(d, "rgb.ply", _rgb_sphere()), [(1, 0, 14), (1, 1, 14), (1, 2, 14), (1, 3, 5), (1, 4, 5), (1, 5, 5)]),
    "rgb_lossless": (lambda d: _ply(d, "rgb.ply", _rgb_sphere()), []),
    # signed and wide integer properties, lossless and quantized from the integer type; per-face integer properties
    "ints_lossless": (lambda d: _ply(d, "ints.ply", _int_sphere()), []),
    "ints_q": (lambda d: _ply(d, "ints.ply", _int_sphere()), [(1, 0, 11), (1, 1, 11), (1, 2, 11), (1, 3, 9), (1, 4, 6), (1, 5, 12), (1, 6, 20), (0, 0, 7)]),
    # (component order after the reader's sorting: x y z red(uchar) material(int, byte offset 13) temp(short, offset 17): misaligned slots)
    "ints_irr_q": (lambda d: _ply(d, "intsi.ply", _int_irregular()), [(1, 0, 10), (1, 1, 10), (1, 2, 10), (1, 3, 5), (1, 4, 13), (1, 5, 7), (0, 0, 9)]),
    "ints_irr_lossless": (lambda d: _ply(d, "intsi.ply", _int_irregular()), []),
}
CONFIG1 = ("sphere35k_lossless", lambda d: _ply(d, "s35k.ply", meshgen.uv_sphere(133, 264)), [])


def _rgb_sphere():
    return meshgen.with_typed_props(meshgen.uv_sphere(31, 49, noise_seed=8), seed=3,
                                    vtx=(("red", np.uint8), ("green", np.uint8), ("blue", np.uint8)))


def _int_sphere():
    return meshgen.with_typed_props(meshgen.uv_sphere(27, 41, noise_seed=12), seed=5,
                                    vtx=(("material", np.int32), ("temp", np.int16), ("flags", np.uint16), ("ident", np.uint32), ("sgn", np.int8)),
                                    face=(("group", np.uint8), ("fid", np.int32)))


def _int_irregular():
    return meshgen.with_typed_props(meshgen.tri_irregular(40, 9), seed=6,
                                    vtx=(("material", np.int32), ("temp", np.int16), ("red", np.uint8)), face=(("group", np.int16),))


def _ply(d, name, mesh):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        meshgen.write_ply(p, mesh)
    return p


def _obj(d, name, multi):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        meshgen.write_obj_latlong(p, 14, 20, multi_region=multi)
    return p


class Case:
    """Everything the parity tests need for one config, all produced by the real reference."""

    def __init__(self, workdir: str, name: str, gen, loq):
        self.name = name
        self.loq = loq
        path = gen(workdir)
        rm = ol.RefMesh(path)
        self.raw = rm.arrays()                       # as read (unquantized), bounds set by the reader
        self.raw_bounds = [(rm.bounds_row(l, 0, la.stride), rm.bounds_row(l, 1, la.stride))
                           for l, la in enumerate(self.raw.lists)]
        for l in range(len(self.raw.lists)):
            rm.set_scale(l)
        self.raw_scale = [rm.bounds_row(l, 2, la.stride) for l, la in enumerate(self.raw.lists)]
        if loq:
            rm.requant(loq)
        rm.traverse()
        self.enc = rm.arrays()                       # quantized + traversal order + final twin table
        self.enc_streams = rm.attr_encode()          # real AttrCoder<Capture> + real model histograms
        hry = os.path.join(workdir, name + ".hry")
        rm.write(hry)
        self.hry_path = hry
        self.src_path = path
        rd = ol.RefMesh(hry)
        self.dec = rd.arrays()                       # decoder-side mesh with the decoded values
        self.dec_streams = rd.logged_streams()       # what the real decoder read from the file
        self.dec_bounds = [(rd.bounds_row(l, 0, la.stride), rd.bounds_row(l, 1, la.stride))
                           for l, la in enumerate(self.dec.lists)]
        if loq:
            rd.requant([], clear=True)
            self.deq = rd.arrays()                   # after requant(clear)
            self.deq_scale = [rd.bounds_row(l, 2, la.stride) for l, la in enumerate(self.dec.lists)]
        else:
            self.deq = None
        rm.close()
        rd.close()

    def decode_input(self) -> capi.MeshArrays:
        m = self.dec.copy()
        m.lists = capi.residual_rows_from_streams(self.dec, self.dec_streams)
        m.emit_types = [ls.type for ls in self.dec_streams.lists]
        return m


_cache = {}


def get_case(workdir: str, name: str) -> Case:
    if name not in _cache:
        if name == CONFIG1[0]:
            _cache[name] = Case(workdir, name, CONFIG1[1], CONFIG1[2])
        else:
            gen, loq = CASES[name]
            _cache[name] = Case(workdir, name, gen, loq)
    return _cache[name]
