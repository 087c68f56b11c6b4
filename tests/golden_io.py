"""(De)serialization of parity cases as small .npz fixtures (tests/golden/).  A loaded fixture has
the same attributes as cases.Case, so the same assertions run against live reference output and
against committed golden vectors."""
from __future__ import annotations

import numpy as np

from harry_b200 import capi

MESH_FIELDS = ("edges", "face_off", "order", "order_f", "vtx_regs", "face_regs", "bind_face", "bind_vtx", "bind_corner",
               "off_reg_face", "off_reg_corner", "off_reg_vtx", "reg_facelist", "reg_cornerlist", "reg_vtxlist")


def _put_mesh(out: dict, pre: str, m: capi.MeshArrays):
    out[pre + "dims"] = np.array([m.nv, m.nf, m.nb_face, m.nb_vtx, m.nb_corner, len(m.lists), 0 if m.order_f is None else 1], dtype=np.int64)
    for f in MESH_FIELDS:
        a = getattr(m, f)
        if a is not None:
            out[pre + f] = np.asarray(a)
    for l, la in enumerate(m.lists):
        out[f"{pre}L{l}_rows"] = la.rows
        out[f"{pre}L{l}_fmt"] = np.array([la.types, la.quants, la.offsets, la.groups if la.groups else [0] * la.ncomp], dtype=np.int64).reshape(4, -1)
        out[f"{pre}L{l}_meta"] = np.array([la.target, la.stride], dtype=np.int64)


def _get_mesh(z, pre: str) -> capi.MeshArrays:
    nv, nf, nbf, nbv, nbc, nl, has_of = (int(x) for x in z[pre + "dims"])
    kw = {f: z[pre + f] for f in MESH_FIELDS if (pre + f) in z}
    if not has_of:
        kw["order_f"] = None
    lists = []
    for l in range(nl):
        fmt = z[f"{pre}L{l}_fmt"]
        target, stride = (int(x) for x in z[f"{pre}L{l}_meta"])
        rows = z[f"{pre}L{l}_rows"].reshape(-1, stride) if stride else np.zeros((z[f"{pre}L{l}_rows"].shape[0], 0), np.uint8)
        lists.append(capi.ListArrays(rows.copy(), [int(x) for x in fmt[0]], [int(x) for x in fmt[1]], [int(x) for x in fmt[2]],
                                     target, [int(x) for x in fmt[3]]))
    return capi.MeshArrays(nv=nv, nf=nf, nb_face=nbf, nb_vtx=nbv, nb_corner=nbc, lists=lists, **kw)


def _put_streams(out: dict, pre: str, s: capi.StreamsPy):
    out[pre + "reg_vtx"] = s.reg_vtx
    out[pre + "reg_face"] = s.reg_face
    out[pre + "n"] = np.array([len(s.lists)], dtype=np.int64)
    for l, ls in enumerate(s.lists):
        for f in ("type", "aux", "symbols", "hist", "type_hist"):
            out[f"{pre}S{l}_{f}"] = getattr(ls, f)


def _get_streams(z, pre: str) -> capi.StreamsPy:
    n = int(z[pre + "n"][0])
    lists = [capi.ListStreamsPy(*(z[f"{pre}S{l}_{f}"] for f in ("type", "aux", "symbols", "hist", "type_hist"))) for l in range(n)]
    return capi.StreamsPy(z[pre + "reg_vtx"], z[pre + "reg_face"], lists)


def save_case(case, path: str):
    out = {}
    _put_mesh(out, "raw_", case.raw)
    _put_mesh(out, "enc_", case.enc)
    _put_mesh(out, "dec_", case.dec)
    _put_streams(out, "es_", case.enc_streams)
    _put_streams(out, "ds_", case.dec_streams)
    for l in range(len(case.raw.lists)):
        out[f"rb{l}_min"], out[f"rb{l}_max"] = case.raw_bounds[l]
        out[f"rb{l}_scale"] = case.raw_scale[l]
        out[f"db{l}_min"], out[f"db{l}_max"] = case.dec_bounds[l]
    out["has_deq"] = np.array([0 if case.deq is None else 1])
    if case.deq is not None:
        _put_mesh(out, "deq_", case.deq)
        for l in range(len(case.dec.lists)):
            out[f"dq{l}_scale"] = case.deq_scale[l]
    out["hry"] = np.frombuffer(open(case.hry_path, "rb").read(), dtype=np.uint8)
    np.savez_compressed(path, **out)


class GoldenCase:
    def __init__(self, path: str):
        z = np.load(path)
        self.raw = _get_mesh(z, "raw_")
        self.enc = _get_mesh(z, "enc_")
        self.dec = _get_mesh(z, "dec_")
        self.enc_streams = _get_streams(z, "es_")
        self.dec_streams = _get_streams(z, "ds_")
        n = len(self.raw.lists)
        self.raw_bounds = [(z[f"rb{l}_min"], z[f"rb{l}_max"]) for l in range(n)]
        self.raw_scale = [z[f"rb{l}_scale"] for l in range(n)]
        self.dec_bounds = [(z[f"db{l}_min"], z[f"db{l}_max"]) for l in range(n)]
        self.deq = None
        if int(z["has_deq"][0]):
            self.deq = _get_mesh(z, "deq_")
            self.deq_scale = [z[f"dq{l}_scale"] for l in range(n)]
        self.hry = z["hry"].tobytes()

    def decode_input(self) -> capi.MeshArrays:
        m = self.dec.copy()
        m.lists = capi.residual_rows_from_streams(self.dec, self.dec_streams)
        m.emit_types = [ls.type for ls in self.dec_streams.lists]
        return m
