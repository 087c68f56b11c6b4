"""Synthetic mesh generators for the five BASELINE.json configs (numpy only).

Every generator is deterministic.  The shapes follow SURVEY.md section 8(d) / Appendix D:

* ``uv_sphere``        configs 1, 2, 5 (UV sphere with poles, optional seeded radial noise)
* ``obj_latlong``      config 3 (lat-long grid OBJ with v / vt / vn, multi-region variant)
* ``poly_grid``        config 4 (mixed tri/quad/5-/6-gon grid + non-manifold fin, per-vertex
                        ``quality`` and per-face ``area`` floats)

The writers emit the binary little-endian PLY / text OBJ the reference's readers parse
(reference: formats/ply/reader.cc:197-432, formats/obj/reader.rl:132-297).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np


# ----------------------------------------------------------------------------------------------
# containers
# ----------------------------------------------------------------------------------------------
@dataclass
class PolyMesh:
    """Indexed polygon mesh: CSR faces + float32 vertex/face property tables."""

    pos: np.ndarray                       # (nv, 3) float32
    face_off: np.ndarray                  # (nf + 1,) uint32 CSR offsets into face_idx
    face_idx: np.ndarray                  # (ne,) uint32 vertex ids
    vtx_props: dict = field(default_factory=dict)    # name -> (nv,) float32 (beyond x, y, z)
    face_props: dict = field(default_factory=dict)   # name -> (nf,) float32

    @property
    def nv(self) -> int:
        return int(self.pos.shape[0])

    @property
    def nf(self) -> int:
        return int(self.face_off.shape[0] - 1)

    @property
    def is_tri(self) -> bool:
        return bool(np.all(np.diff(self.face_off.astype(np.int64)) == 3))


# ----------------------------------------------------------------------------------------------
# config 1 / 2 / 5: UV sphere
# ----------------------------------------------------------------------------------------------
def uv_sphere(nr: int, ns: int, noise_seed: int | None = None, noise_sigma: float = 0.01) -> PolyMesh:
    """UV sphere with poles: vertex 0 = north pole, rings i = 1..nr-1 of ``ns`` vertices, south
    pole last.  ``nr=133, ns=264`` gives 34 850 vertices / 69 696 triangles (config 1);
    ``nr=2237, ns=4472`` gives 9 999 394 / 19 998 784 (config 2); ``nr=225, ns=447`` gives
    100 130 / 200 256 (config 5).  ``noise_seed`` adds radial noise r = 1 + sigma * N(0, 1)."""
    assert nr >= 2 and ns >= 3
    nv = 2 + (nr - 1) * ns
    i = np.arange(1, nr, dtype=np.float64)
    j = np.arange(ns, dtype=np.float64)
    theta = np.pi * i / nr
    phi = 2.0 * np.pi * j / ns
    st, ct = np.sin(theta), np.cos(theta)
    ring = np.empty((nr - 1, ns, 3), dtype=np.float64)
    ring[:, :, 0] = st[:, None] * np.cos(phi)[None, :]
    ring[:, :, 1] = st[:, None] * np.sin(phi)[None, :]
    ring[:, :, 2] = ct[:, None]
    pos = np.empty((nv, 3), dtype=np.float64)
    pos[0] = (0.0, 0.0, 1.0)
    pos[1:-1] = ring.reshape(-1, 3)
    pos[-1] = (0.0, 0.0, -1.0)
    if noise_seed is not None:
        rng = np.random.default_rng(noise_seed)
        pos *= (1.0 + noise_sigma * rng.standard_normal(nv))[:, None]
    pos = pos.astype(np.float32)

    def idx(ii, jj):
        return 1 + (ii - 1) * ns + (jj % ns)

    jj = np.arange(ns, dtype=np.int64)
    # north cap
    top = np.stack([np.zeros(ns, dtype=np.int64), idx(1, jj), idx(1, jj + 1)], axis=1)
    # body quads -> two triangles (a,b,c),(a,c,d)
    ii = np.arange(1, nr - 1, dtype=np.int64)
    I, J = np.meshgrid(ii, jj, indexing="ij")
    a, b, c, d = idx(I, J), idx(I + 1, J), idx(I + 1, J + 1), idx(I, J + 1)
    body = np.stack([np.stack([a, b, c], -1), np.stack([a, c, d], -1)], axis=2).reshape(-1, 3)
    last = nv - 1
    bot = np.stack([np.full(ns, last, dtype=np.int64), idx(nr - 1, jj + 1), idx(nr - 1, jj)], axis=1)
    tris = np.concatenate([top, body, bot], axis=0).astype(np.uint32)
    nf = tris.shape[0]
    face_off = (np.arange(nf + 1, dtype=np.uint64) * 3).astype(np.uint32)
    return PolyMesh(pos=pos, face_off=face_off, face_idx=tris.reshape(-1))


# ----------------------------------------------------------------------------------------------
# config 4: mixed polygon grid with a non-manifold fin
# ----------------------------------------------------------------------------------------------
def poly_grid(n: int = 60, seed: int = 7) -> PolyMesh:
    """(n+1)^2 grid whose cells are grouped by k = (7x + 3y) mod 5 into hexagons, pentagon +
    triangle, two triangles or a quad, plus a two-triangle fin on edge (1,0)-(0,0).  Per-vertex
    float ``quality`` and per-face float ``area`` come from ``default_rng(seed).random()``.
    No T-junction vertices are produced (the reference does not round-trip those)."""
    def vid(x, y):
        return y * (n + 1) + x

    pos = []
    for y in range(n + 1):
        for x in range(n + 1):
            pos.append((float(x), float(y), 0.1 * np.sin(0.3 * x) * np.cos(0.2 * y)))
    faces = []
    for y in range(n):
        x = 0
        while x < n:
            k = (7 * x + 3 * y) % 5
            a, b, c, d = vid(x, y), vid(x + 1, y), vid(x + 1, y + 1), vid(x, y + 1)
            if k == 0 and x + 1 < n:
                b2, c2 = vid(x + 2, y), vid(x + 2, y + 1)
                faces.append((a, b, b2, c2, c, d))
                x += 2
            elif k == 1 and x + 1 < n:
                b2, c2 = vid(x + 2, y), vid(x + 2, y + 1)
                faces.append((a, b, b2, c, d))
                faces.append((b2, c2, c))
                x += 2
            elif k == 2:
                faces.append((a, b, c))
                faces.append((a, c, d))
                x += 1
            else:
                faces.append((a, b, c, d))
                x += 1
    f1 = len(pos)
    pos.append((0.5, -0.5, 1.0))
    f2 = len(pos)
    pos.append((0.5, -0.5, -1.0))
    faces.append((vid(1, 0), vid(0, 0), f1))
    faces.append((vid(1, 0), vid(0, 0), f2))
    pos = np.asarray(pos, dtype=np.float32)
    deg = np.asarray([len(f) for f in faces], dtype=np.uint32)
    face_off = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    face_idx = np.concatenate([np.asarray(f, dtype=np.uint32) for f in faces])
    rng = np.random.default_rng(seed)
    quality = rng.random(pos.shape[0]).astype(np.float32)
    area = rng.random(len(faces)).astype(np.float32)
    return PolyMesh(pos=pos, face_off=face_off, face_idx=face_idx,
                    vtx_props={"quality": quality}, face_props={"area": area})


def tri_irregular(n: int = 48, seed: int = 11, holes: bool = True) -> PolyMesh:
    """Irregular triangulation: a jittered (n+1)^2 height field, every cell split along a random
    diagonal (vertex valences 3..8, vertices with 1, 2, 3 and more parallelograms), a few cells
    removed (interior borders), rough z so that residuals are large and hit the escape branch of
    the residual mapping near the range limits.  Exercises the irregular paths of the vertex decoder."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:n + 1, 0:n + 1]
    pos = np.empty(((n + 1) * (n + 1), 3), dtype=np.float64)
    pos[:, 0] = xs.ravel() + 0.35 * rng.standard_normal(xs.size) * ((xs.ravel() > 0) & (xs.ravel() < n))
    pos[:, 1] = ys.ravel() + 0.35 * rng.standard_normal(ys.size) * ((ys.ravel() > 0) & (ys.ravel() < n))
    pos[:, 2] = np.sin(0.4 * xs.ravel()) * np.cos(0.3 * ys.ravel()) + 0.3 * rng.standard_normal(xs.size)
    # a band of exactly flat, exactly extreme z (escape / saturation regimes, constant runs)
    pos[(ys.ravel() % 7) == 3, 2] = pos[:, 2].min()

    def vid(x, y):
        return y * (n + 1) + x

    flip = rng.random((n, n)) < 0.5
    drop = (rng.random((n, n)) < 0.02) if holes else np.zeros((n, n), dtype=bool)
    faces = []
    for y in range(n):
        for x in range(n):
            if drop[y, x]:
                continue
            a, b, c, d = vid(x, y), vid(x + 1, y), vid(x + 1, y + 1), vid(x, y + 1)
            if flip[y, x]:
                faces.append((a, b, d))
                faces.append((b, c, d))
            else:
                faces.append((a, b, c))
                faces.append((a, c, d))
    faces = np.asarray(faces, dtype=np.uint32)
    face_off = (3 * np.arange(faces.shape[0] + 1)).astype(np.uint32)
    return PolyMesh(pos=pos.astype(np.float32), face_off=face_off, face_idx=faces.ravel())


def cones(k: int = 80, valence: int = 90, seed: int = 2, open_every: int = 0) -> PolyMesh:
    """k separate cones: an apex joined to a closed ring of `valence` vertices.  Every apex is a wide
    fan for the fan gather (more than 64 steps); more cones than the wide path takes (64) make the
    surplus fall back to one-thread walks.  Noisy positions.  open_every = n > 0: every n-th cone
    lacks three of its triangles, so its apex sits on a border (an OPEN wide fan: the walk runs
    forward from the gate to the border, then backward from the gate's predecessor)."""
    rng = np.random.default_rng(seed)
    pos, faces = [], []
    for c in range(k):
        base = len(pos)
        cx, cy = 3.0 * (c % 10), 3.0 * (c // 10)
        pos.append((cx, cy, 1.0 + 0.1 * rng.standard_normal()))
        for j in range(valence):
            a = 2.0 * np.pi * j / valence
            r = 1.0 + 0.05 * rng.standard_normal()
            pos.append((cx + r * np.cos(a), cy + r * np.sin(a), 0.05 * rng.standard_normal()))
        gap = open_every > 0 and c % open_every == 0
        for j in range(valence):
            if gap and valence // 3 <= j < valence // 3 + 3:
                continue
            faces.append((base, base + 1 + j, base + 1 + (j + 1) % valence))
    faces = np.asarray(faces, dtype=np.uint32)
    face_off = (3 * np.arange(faces.shape[0] + 1)).astype(np.uint32)
    return PolyMesh(pos=np.asarray(pos, dtype=np.float32), face_off=face_off, face_idx=faces.ravel())


def soup(nv: int = 30, nf: int = 400, seed: int = 1, min_deg: int = 1, max_deg: int = 6) -> PolyMesh:
    """Random polygon soup over few vertices: heavily non-manifold (many half-edges per directed edge), with
    degenerate edges (a, a) and faces of 1 or 2 corners when min_deg allows.  Input for the twin matching only
    (structs/conn.h:178-190: which half-edges pair up depends on their file order)."""
    rng = np.random.default_rng(seed)
    deg = rng.integers(min_deg, max_deg + 1, nf)
    face_off = np.concatenate([[0], np.cumsum(deg)]).astype(np.uint32)
    face_idx = rng.integers(0, nv, int(face_off[-1])).astype(np.uint32)
    face_idx[rng.integers(0, face_idx.shape[0])] = nv - 1   # the reader derives nv from the largest index
    return PolyMesh(pos=rng.random((nv, 3)).astype(np.float32), face_off=face_off, face_idx=face_idx)


# ----------------------------------------------------------------------------------------------
# PLY writer (binary little endian)
# ----------------------------------------------------------------------------------------------
PLY_TYPE = {"float32": "float", "float64": "double", "uint8": "uchar", "int8": "char", "uint16": "ushort", "int16": "short",
            "uint32": "uint", "int32": "int"}


def with_typed_props(m: PolyMesh, seed: int = 1, vtx: tuple = (), face: tuple = ()) -> PolyMesh:
    """Adds typed per-vertex / per-face properties (PLY scalar types other than float: SURVEY Appendix C.13, the
    `uchar` colours next to float coordinates of scanned meshes).  ``vtx`` / ``face``: (name, numpy dtype) pairs.
    Values are smooth functions of the position plus seeded noise, spread over most of the type's range, with
    negative values for the signed types."""
    rng = np.random.default_rng(seed)
    nv, nf = m.nv, m.nf
    deg = np.diff(m.face_off.astype(np.int64))
    centre = np.add.reduceat(m.pos[m.face_idx.astype(np.int64)], m.face_off[:-1].astype(np.int64), axis=0) / deg[:, None]

    def make(points, k, dt):
        dt = np.dtype(dt)
        base = 0.5 + 0.45 * np.sin(3.0 * points[:, k % 3] + 0.7 * k) * np.cos(2.0 * points[:, (k + 1) % 3])
        base = base + 0.02 * rng.standard_normal(points.shape[0])
        if dt.kind == "f":
            return base.astype(dt)
        info = np.iinfo(dt)
        lo, hi = info.min, info.max
        if dt.itemsize >= 4:
            # int32 arithmetic of the reference is not promoted: a difference of two values must not overflow (signed
            # overflow is undefined there, prediction.h:82-99), so signed 32-bit values stay within +-2^28
            lo, hi = (-(1 << 28), 1 << 28) if dt.kind == "i" else (0, info.max // 2)
        return np.clip(np.rint(lo + np.clip(base, 0.0, 1.0) * (float(hi) - float(lo))), info.min, info.max).astype(dt)

    out = PolyMesh(pos=m.pos, face_off=m.face_off, face_idx=m.face_idx, vtx_props=dict(m.vtx_props), face_props=dict(m.face_props))
    for k, (name, dt) in enumerate(vtx):
        out.vtx_props[name] = make(m.pos, k, dt)
    for k, (name, dt) in enumerate(face):
        out.face_props[name] = make(centre, k + 5, dt)
    return out


def write_ply(path: str, m: PolyMesh) -> None:
    """Binary little-endian PLY; every property keeps the numpy dtype of its array (float32 unless made otherwise)."""
    nv, nf = m.nv, m.nf
    vprops = [("x", m.pos[:, 0]), ("y", m.pos[:, 1]), ("z", m.pos[:, 2])] + [(nm, np.asarray(a)) for nm, a in m.vtx_props.items()]
    vprops = [(nm, a if a.dtype.name in PLY_TYPE else a.astype(np.float32)) for nm, a in vprops]
    fprops = [(nm, np.asarray(a)) for nm, a in m.face_props.items()]
    fprops = [(nm, a if a.dtype.name in PLY_TYPE else a.astype(np.float32)) for nm, a in fprops]
    hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {nv}"]
    hdr += [f"property {PLY_TYPE[a.dtype.name]} {nm}" for nm, a in vprops]
    hdr += [f"element face {nf}", "property list uchar int vertex_indices"]
    hdr += [f"property {PLY_TYPE[a.dtype.name]} {nm}" for nm, a in fprops]
    hdr += ["end_header"]
    vt = np.empty(nv, dtype=[(nm, a.dtype.newbyteorder("<")) for nm, a in vprops])
    for nm, a in vprops:
        vt[nm] = a
    with open(path, "wb") as f:
        f.write(("\n".join(hdr) + "\n").encode("ascii"))
        f.write(vt.tobytes())
        if m.is_tri:
            rec = np.empty(nf, dtype=[("n", "u1"), ("i", "<i4", 3)] + [(nm, a.dtype.newbyteorder("<")) for nm, a in fprops])
            rec["n"] = 3
            rec["i"] = m.face_idx.reshape(-1, 3).astype("<i4")
            for nm, a in fprops:
                rec[nm] = a
            f.write(rec.tobytes())
        else:
            out = bytearray()
            for fi in range(nf):
                s, e = int(m.face_off[fi]), int(m.face_off[fi + 1])
                out += bytes([e - s])
                out += m.face_idx[s:e].astype("<i4").tobytes()
                for _, fp in fprops:
                    out += fp[fi:fi + 1].astype(fp.dtype.newbyteorder("<")).tobytes()
            f.write(bytes(out))


# ----------------------------------------------------------------------------------------------
# config 3: OBJ lat-long sphere with v / vt / vn (+ multi-region variant)
# ----------------------------------------------------------------------------------------------
def write_obj_latlong(path: str, nr: int = 40, ns: int = 60, multi_region: bool = False) -> dict:
    """(nr+1) x ns lat-long grid without poles, ``v = vn = (sin t cos p, sin t sin p, cos t)``,
    ``vt = (j/ns, i/nr)``; all ``v`` lines, then all ``vt``, then all ``vn`` so the lists are
    numbered v -> 0, vt -> 1, vn -> 2 (formats/obj/reader.rl:132-147).  ``multi_region`` adds a
    second material, a band of ``f v//vn`` faces, a band of bare ``f v`` faces and a band of
    6-coordinate vertices (a second VTX list / region, reader.rl:141-142,196-200)."""
    nrow = nr + 1
    lines = []
    if multi_region:
        mtl = os.path.splitext(path)[0] + ".mtl"
        with open(mtl, "w") as f:
            f.write("newmtl matA\nKd 0.800000 0.100000 0.100000\nnewmtl matB\nKd 0.100000 0.200000 0.900000\n")
        lines.append(f"mtllib {os.path.basename(mtl)}")
    colour_rows = set(range(nrow // 2, nrow // 2 + 3)) if multi_region else set()
    vs, vts, vns = [], [], []
    for i in range(nrow):
        t = np.pi * (i + 0.5) / (nr + 1)
        for j in range(ns):
            p = 2.0 * np.pi * j / ns
            x, y, z = np.sin(t) * np.cos(p), np.sin(t) * np.sin(p), np.cos(t)
            if i in colour_rows:
                vs.append("v %.6f %.6f %.6f %.6f %.6f %.6f" % (x, y, z, j / ns, i / nr, 0.5))
            else:
                vs.append("v %.6f %.6f %.6f" % (x, y, z))
            vts.append("vt %.6f %.6f" % (j / ns, i / nr))
            vns.append("vn %.6f %.6f %.6f" % (x, y, z))
    lines += vs + vts + vns

    def idx(i, j):
        return i * ns + (j % ns) + 1

    ntri = 0
    cur_mtl = None
    for i in range(nr):
        if multi_region:
            want = "matA" if i < nr // 2 else "matB"
            if want != cur_mtl:
                lines.append(f"usemtl {want}")
                cur_mtl = want
        for j in range(ns):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            for tri in ((a, b, c), (a, c, d)):
                if multi_region and i in (3, 4):
                    lines.append("f " + " ".join(f"{v}//{v}" for v in tri))
                elif multi_region and i in (7, 8):
                    lines.append("f " + " ".join(f"{v}" for v in tri))
                else:
                    lines.append("f " + " ".join(f"{v}/{v}/{v}" for v in tri))
                ntri += 1
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return {"nv": nrow * ns, "nf": ntri}
