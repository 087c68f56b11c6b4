"""Flatten a PolyMesh into the hb_mesh_desc arrays WITHOUT the reference's host code.

The product pipeline gets its traversal order and final twin table from the Cut-Border-Machine
traversal of the reference host code (out of scope of the GPU path).  This module
provides a self-contained stand-in for tests, smoke() and micro-benchmarks: twins are matched the
way mesh::conn::Builder::add_edge does for manifold input (structs/conn.h:201-214, first directed
edge wins, later duplicates stay borders), and the traversal order is "first appearance in the
face list" -- a valid order for the attribute coder (any order with org(order[i]) distinct works),
although not the CBM's.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .meshgen import PolyMesh


def build_edges(face_off: np.ndarray, face_idx: np.ndarray) -> np.ndarray:
    """(ne, 3) uint32 records {org, twin_face, twin_edge}; unmatched half-edges are their own twin."""
    ne = int(face_idx.shape[0])
    nf = int(face_off.shape[0] - 1)
    deg = np.diff(face_off.astype(np.int64))
    hface = np.repeat(np.arange(nf, dtype=np.int64), deg)
    hloc = np.arange(ne, dtype=np.int64) - face_off[:-1].astype(np.int64)[hface]
    nxt = np.arange(ne, dtype=np.int64) + 1
    last = hloc == deg[hface] - 1
    nxt[last] = face_off[:-1].astype(np.int64)[hface[last]]
    a = face_idx.astype(np.int64)
    b = a[nxt]
    twin = np.arange(ne, dtype=np.int64)
    nv = int(a.max()) + 1 if ne else 0
    key = a * nv + b
    rev = b * nv + a
    order = np.argsort(key, kind="stable")
    skey = key[order]
    if ne == 0 or np.all(skey[1:] != skey[:-1]):
        # every directed edge is unique (manifold orientation): vectorized matching
        pos = np.searchsorted(skey, rev)
        pos_c = np.minimum(pos, ne - 1)
        hit = skey[pos_c] == rev
        twin[hit] = order[pos_c[hit]]
    else:
        # duplicates: sequential first-come matching (exact for any input, O(ne) dict operations)
        pending = {}
        al, bl = a.tolist(), b.tolist()
        for h in range(ne):
            t = pending.pop((bl[h], al[h]), None)
            if t is not None:
                twin[h] = t
                twin[t] = h
            else:
                pending.setdefault((al[h], bl[h]), h)
    edges = np.empty((ne, 3), dtype=np.uint32)
    edges[:, 0] = face_idx
    edges[:, 1] = hface[twin]
    edges[:, 2] = hloc[twin]
    return edges


def first_appearance_order(face_off: np.ndarray, face_idx: np.ndarray):
    """order[i] = (face, local edge) of the first half-edge whose origin is the i-th distinct vertex
    in face-list order."""
    nf = int(face_off.shape[0] - 1)
    deg = np.diff(face_off.astype(np.int64))
    hface = np.repeat(np.arange(nf, dtype=np.int64), deg)
    hloc = np.arange(face_idx.shape[0], dtype=np.int64) - face_off[:-1].astype(np.int64)[hface]
    _, first = np.unique(face_idx, return_index=True)
    first = np.sort(first)
    order = np.stack([hface[first], hloc[first]], axis=1).astype(np.uint32)
    return order


def mesh_arrays(pm: PolyMesh, order: np.ndarray | None = None, order_f: np.ndarray | None = None) -> capi.MeshArrays:
    """PLY-style flattening (formats/ply/reader.cc:388-412): list 0 = face properties (FACE),
    list 1 = vertex properties (VTX), one region each, identity bindings."""
    nv, nf = pm.nv, pm.nf
    edges = build_edges(pm.face_off, pm.face_idx)
    if order is None:
        order = first_appearance_order(pm.face_off, pm.face_idx)
    if order_f is None:
        order_f = np.stack([np.arange(nf, dtype=np.uint32), np.zeros(nf, dtype=np.uint32)], axis=1)
    vcols = [pm.pos[:, 0], pm.pos[:, 1], pm.pos[:, 2]] + [v for v in pm.vtx_props.values()]
    vgroups = [0, 0, 0] + [3 + k for k in range(len(pm.vtx_props))]
    vlist = capi.make_list(vcols, [capi.FLOAT] * len(vcols), capi.T_VTX, vgroups)
    if pm.face_props:
        fcols = [v for v in pm.face_props.values()]
        flist = capi.make_list(fcols, [capi.FLOAT] * len(fcols), capi.T_FACE)
    else:
        flist = capi.empty_list(nf, capi.T_FACE)
    return capi.MeshArrays(
        nv=nv, nf=nf, edges=edges, face_off=pm.face_off.astype(np.uint32), order=order, order_f=order_f,
        vtx_regs=np.zeros(nv, np.uint16), face_regs=np.zeros(nf, np.uint16),
        nb_face=1, nb_vtx=1, nb_corner=0,
        bind_face=np.arange(nf, dtype=np.uint32), bind_vtx=np.arange(nv, dtype=np.uint32),
        bind_corner=np.zeros(0, np.uint32),
        off_reg_face=np.array([0, 1], np.int32), off_reg_corner=np.array([0, 0], np.int32),
        off_reg_vtx=np.array([0, 1], np.int32),
        reg_facelist=np.array([0], np.uint16), reg_cornerlist=np.zeros(0, np.uint16),
        reg_vtxlist=np.array([1], np.uint16), lists=[flist, vlist])
