"""ctypes mirror of include/harry_b200.h plus numpy containers for its structs.

The product library is ``harry_b200/libharry_b200.so`` (CUDA, sm_100a, built by
``__graft_entry__.build()``).  There is no CPU fallback: ``load_library()`` raises if the
library is missing, and every ``hb_*`` entry point fails with HB_ERR_CUDA without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

HB_MAX_COMP = 32

# mixing::Type (structs/mixing.h:19)
FLOAT, DOUBLE, ULONG, LONG, UINT, INT, USHORT, SHORT, UCHAR, CHAR, TYPE_NONE = range(11)
TYPE_SIZE = (4, 8, 8, 8, 4, 4, 2, 2, 1, 1, 0)
TYPE_NP = ("<f4", "<f8", "<u8", "<i8", "<u4", "<i4", "<u2", "<i2", "u1", "i1", None)
# mesh::attr::Target (structs/attr.h:22)
T_FACE, T_VTX, T_CORNER, T_NONE = range(4)
# hry::AttrType (formats/hry/models.h:21)
DATA, HIST, LHIST = range(3)


def quant_stype(q: int) -> int:
    """Fmt::quant_type, structs/mixing.h:101-108."""
    if q <= 8:
        return UCHAR
    if q <= 16:
        return USHORT
    if q <= 32:
        return UINT
    return ULONG


class ListDesc(C.Structure):
    _fields_ = [
        ("rows", C.c_void_p),
        ("nrows", C.c_uint32),
        ("stride", C.c_uint32),
        ("ncomp", C.c_uint16),
        ("target", C.c_uint8),
        ("reserved", C.c_uint8),
        ("type", C.c_uint8 * HB_MAX_COMP),
        ("quant", C.c_uint8 * HB_MAX_COMP),
        ("offset", C.c_uint16 * HB_MAX_COMP),
    ]


class MeshDesc(C.Structure):
    _fields_ = [
        ("nv", C.c_uint32), ("nf", C.c_uint32), ("ne", C.c_uint32),
        ("edges", C.c_void_p),
        ("face_off", C.c_void_p),
        ("order", C.c_void_p),
        ("norder", C.c_uint32),
        ("order_f", C.c_void_p),
        ("norder_f", C.c_uint32),
        ("vtx_regs", C.c_void_p),
        ("face_regs", C.c_void_p),
        ("nb_face", C.c_uint16), ("nb_vtx", C.c_uint16), ("nb_corner", C.c_uint16),
        ("nregs_face", C.c_uint16), ("nregs_vtx", C.c_uint16),
        ("nlists", C.c_uint16),
        ("bind_face_attr", C.c_void_p),
        ("bind_vtx_attr", C.c_void_p),
        ("bind_corner_attr", C.c_void_p),
        ("off_reg_face", C.c_void_p),
        ("off_reg_corner", C.c_void_p),
        ("off_reg_vtx", C.c_void_p),
        ("reg_facelist", C.c_void_p),
        ("reg_cornerlist", C.c_void_p),
        ("reg_vtxlist", C.c_void_p),
        ("lists", C.POINTER(ListDesc)),
        ("emit_type", C.POINTER(C.c_void_p)),
        ("emit_count", C.POINTER(C.c_uint32)),
    ]


class ListStreams(C.Structure):
    _fields_ = [
        ("n_emit", C.c_uint32),
        ("n_data", C.c_uint32),
        ("sym_stride", C.c_uint32),
        ("reserved", C.c_uint32),
        ("type", C.POINTER(C.c_uint8)),
        ("aux", C.POINTER(C.c_uint32)),
        ("symbols", C.POINTER(C.c_uint8)),
        ("hist", C.POINTER(C.c_uint64)),
        ("type_hist", C.c_uint64 * 4),
    ]


class Streams(C.Structure):
    _fields_ = [
        ("n_vtx", C.c_uint32),
        ("n_face", C.c_uint32),
        ("reg_vtx", C.POINTER(C.c_uint16)),
        ("reg_face", C.POINTER(C.c_uint16)),
        ("nlists", C.c_uint16),
        ("lists", C.POINTER(ListStreams)),
    ]


class BatchStreams(C.Structure):
    _fields_ = [("n", C.c_uint32), ("mesh", C.POINTER(Streams)), ("priv", C.c_void_p)]


class QuantReq(C.Structure):
    _fields_ = [("list", C.c_uint32), ("new_quant", C.c_uint8 * HB_MAX_COMP), ("groups", C.c_uint8 * HB_MAX_COMP)]


class DequantReq(C.Structure):
    _fields_ = [("list", C.c_uint32), ("bounds", C.c_void_p)]


# ----------------------------------------------------------------------------------------------
# numpy containers
# ----------------------------------------------------------------------------------------------
@dataclass
class ListArrays:
    """One attribute list: AoS rows (nrows, stride) uint8 + the mixing::Fmt description."""
    rows: np.ndarray
    types: list
    quants: list
    offsets: list
    target: int
    groups: list = field(default_factory=list)   # interpretation-group leader per component

    @property
    def nrows(self) -> int:
        return int(self.rows.shape[0])

    @property
    def stride(self) -> int:
        return int(self.rows.shape[1]) if self.rows.ndim == 2 else 0

    @property
    def ncomp(self) -> int:
        return len(self.types)

    def stype(self, j: int) -> int:
        return quant_stype(self.quants[j]) if self.quants[j] else self.types[j]

    @property
    def sym_stride(self) -> int:
        return sum(TYPE_SIZE[self.stype(j)] for j in range(self.ncomp))

    def copy(self) -> "ListArrays":
        return ListArrays(self.rows.copy(), list(self.types), list(self.quants), list(self.offsets),
                          self.target, list(self.groups))

    def component(self, j: int, stype: int | None = None) -> np.ndarray:
        """Typed strided view of component j (in its storage type by default)."""
        st = self.stype(j) if stype is None else stype
        sz = TYPE_SIZE[st]
        off = self.offsets[j]
        if self.nrows == 0:
            return np.zeros(0, dtype=TYPE_NP[st])
        return np.ndarray((self.nrows,), dtype=TYPE_NP[st], buffer=self.rows.data, offset=off,
                          strides=(self.stride,)) if sz else np.zeros(0)

    def to_desc(self) -> ListDesc:
        d = ListDesc()
        self.rows = np.ascontiguousarray(self.rows, dtype=np.uint8)
        d.rows = self.rows.ctypes.data if self.rows.size else None
        d.nrows = self.nrows
        d.stride = self.stride
        d.ncomp = self.ncomp
        d.target = self.target
        for j in range(self.ncomp):
            d.type[j] = self.types[j]
            d.quant[j] = self.quants[j]
            d.offset[j] = self.offsets[j]
        return d

    def sync_from_desc(self, d: ListDesc) -> None:
        self.quants = [int(d.quant[j]) for j in range(self.ncomp)]


def make_list(cols: list, types: list, target: int, groups: list | None = None) -> ListArrays:
    """Pack typed columns into AoS rows the way mixing::Fmt::add lays them out (mixing.h:54-61)."""
    n = len(cols[0]) if cols else 0
    offsets, off = [], 0
    for t in types:
        offsets.append(off)
        off += TYPE_SIZE[t]
    rows = np.zeros((n, off), dtype=np.uint8)
    la = ListArrays(rows, list(types), [0] * len(types), offsets, target,
                    list(groups) if groups is not None else list(range(len(types))))
    for j, col in enumerate(cols):
        la.component(j)[:] = np.asarray(col, dtype=TYPE_NP[types[j]])
    return la


def empty_list(nrows: int, target: int) -> ListArrays:
    return ListArrays(np.zeros((nrows, 0), dtype=np.uint8), [], [], [], target, [])


@dataclass
class MeshArrays:
    """Flattened mesh as the attribute coder sees it (fields of hb_mesh_desc)."""
    nv: int
    nf: int
    edges: np.ndarray            # (ne, 3) uint32: org, twin_face, twin_edge (low 16 bits)
    face_off: np.ndarray         # (nf + 1,) uint32
    order: np.ndarray            # (norder, 2) uint32: face, edge
    order_f: np.ndarray | None   # (norder_f, 2) uint32 or None (decoder order)
    vtx_regs: np.ndarray         # (nv,) uint16
    face_regs: np.ndarray        # (nf,) uint16
    nb_face: int
    nb_vtx: int
    nb_corner: int
    bind_face: np.ndarray        # (nf * nb_face,) uint32
    bind_vtx: np.ndarray
    bind_corner: np.ndarray
    off_reg_face: np.ndarray     # int32
    off_reg_corner: np.ndarray
    off_reg_vtx: np.ndarray
    reg_facelist: np.ndarray     # uint16
    reg_cornerlist: np.ndarray
    reg_vtxlist: np.ndarray
    lists: list                  # [ListArrays]
    emit_types: list | None = None   # decode only: per list uint8 type stream (or None)
    _keep: list = field(default_factory=list, repr=False)

    @property
    def ne(self) -> int:
        return int(self.edges.shape[0])

    def copy(self) -> "MeshArrays":
        import copy as _c
        m = _c.copy(self)
        m.lists = [l.copy() for l in self.lists]
        m._keep = []
        return m

    def n_attrs(self) -> int:
        """Vertex-attributes of the hot path: sum over lists of rows x components."""
        return int(sum(l.nrows * l.ncomp for l in self.lists))

    def to_desc(self) -> MeshDesc:
        def prep(name, dtype):
            a = getattr(self, name)
            if a is None:
                return None
            a = np.ascontiguousarray(a, dtype=dtype)
            setattr(self, name, a)
            return a.ctypes.data if a.size else None

        d = MeshDesc()
        d.nv, d.nf, d.ne = self.nv, self.nf, self.ne
        d.edges = prep("edges", np.uint32)
        d.face_off = prep("face_off", np.uint32)
        d.order = prep("order", np.uint32)
        d.norder = int(self.order.shape[0])
        d.order_f = prep("order_f", np.uint32)
        d.norder_f = 0 if self.order_f is None else int(self.order_f.shape[0])
        d.vtx_regs = prep("vtx_regs", np.uint16)
        d.face_regs = prep("face_regs", np.uint16)
        d.nb_face, d.nb_vtx, d.nb_corner = self.nb_face, self.nb_vtx, self.nb_corner
        d.nregs_face = len(self.off_reg_face) - 1
        d.nregs_vtx = len(self.off_reg_vtx) - 1
        d.nlists = len(self.lists)
        d.bind_face_attr = prep("bind_face", np.uint32)
        d.bind_vtx_attr = prep("bind_vtx", np.uint32)
        d.bind_corner_attr = prep("bind_corner", np.uint32)
        d.off_reg_face = prep("off_reg_face", np.int32)
        d.off_reg_corner = prep("off_reg_corner", np.int32)
        d.off_reg_vtx = prep("off_reg_vtx", np.int32)
        d.reg_facelist = prep("reg_facelist", np.uint16)
        d.reg_cornerlist = prep("reg_cornerlist", np.uint16)
        d.reg_vtxlist = prep("reg_vtxlist", np.uint16)
        arr = (ListDesc * max(1, len(self.lists)))()
        for i, l in enumerate(self.lists):
            arr[i] = l.to_desc()
        d.lists = arr
        self._keep = [arr]
        if self.emit_types is not None:
            tarr = (C.c_void_p * max(1, len(self.lists)))()
            carr = (C.c_uint32 * max(1, len(self.lists)))()
            held = []
            for i in range(len(self.lists)):
                t = self.emit_types[i] if i < len(self.emit_types) else None
                if t is None or len(t) == 0:
                    tarr[i] = None
                else:
                    t = np.ascontiguousarray(t, dtype=np.uint8)
                    held.append(t)
                    tarr[i] = t.ctypes.data
                    carr[i] = len(t)
            d.emit_type = tarr
            d.emit_count = carr
            self._keep += [tarr, carr, held]
        return d


def mesh_from_desc(d: MeshDesc, groups_fn=None) -> MeshArrays:
    """Deep copy of a C hb_mesh_desc (e.g. one filled by the reference harness) into numpy."""
    def arr(ptr, n, dtype):
        if not ptr or n == 0:
            return np.zeros(0, dtype=dtype)
        buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(ptr)
        return np.frombuffer(buf, dtype=dtype).copy()

    nv, nf, ne = d.nv, d.nf, d.ne
    edges_raw = arr(d.edges, ne * 3, np.uint32).reshape(-1, 3)
    edges = edges_raw.copy()
    edges[:, 2] &= 0xFFFF
    order = arr(d.order, d.norder * 2, np.uint32).reshape(-1, 2).copy()
    order[:, 1] &= 0xFFFF
    order_f = None
    if d.order_f:
        order_f = arr(d.order_f, d.norder_f * 2, np.uint32).reshape(-1, 2).copy()
        order_f[:, 1] &= 0xFFFF
    off_f = arr(d.off_reg_face, d.nregs_face + 1, np.int32)
    off_c = arr(d.off_reg_corner, d.nregs_face + 1, np.int32)
    off_v = arr(d.off_reg_vtx, d.nregs_vtx + 1, np.int32)
    lists = []
    for i in range(d.nlists):
        L = d.lists[i]
        rows = arr(L.rows, L.nrows * L.stride, np.uint8).reshape(L.nrows, L.stride)
        la = ListArrays(rows, [int(L.type[j]) for j in range(L.ncomp)], [int(L.quant[j]) for j in range(L.ncomp)],
                        [int(L.offset[j]) for j in range(L.ncomp)], int(L.target),
                        list(groups_fn(i)) if groups_fn else list(range(L.ncomp)))
        lists.append(la)
    return MeshArrays(
        nv=nv, nf=nf, edges=edges, face_off=arr(d.face_off, nf + 1, np.uint32), order=order, order_f=order_f,
        vtx_regs=arr(d.vtx_regs, nv, np.uint16), face_regs=arr(d.face_regs, nf, np.uint16),
        nb_face=d.nb_face, nb_vtx=d.nb_vtx, nb_corner=d.nb_corner,
        bind_face=arr(d.bind_face_attr, nf * d.nb_face, np.uint32),
        bind_vtx=arr(d.bind_vtx_attr, nv * d.nb_vtx, np.uint32),
        bind_corner=arr(d.bind_corner_attr, ne * d.nb_corner, np.uint32),
        off_reg_face=off_f, off_reg_corner=off_c, off_reg_vtx=off_v,
        reg_facelist=arr(d.reg_facelist, int(off_f[-1]) if len(off_f) else 0, np.uint16),
        reg_cornerlist=arr(d.reg_cornerlist, int(off_c[-1]) if len(off_c) else 0, np.uint16),
        reg_vtxlist=arr(d.reg_vtxlist, int(off_v[-1]) if len(off_v) else 0, np.uint16),
        lists=lists)


@dataclass
class ListStreamsPy:
    type: np.ndarray
    aux: np.ndarray
    symbols: np.ndarray      # (n_data, sym_stride) uint8
    hist: np.ndarray         # (sym_stride, 256) uint64
    type_hist: np.ndarray


@dataclass
class StreamsPy:
    reg_vtx: np.ndarray
    reg_face: np.ndarray
    lists: list

    def equal(self, o: "StreamsPy") -> tuple:
        if not np.array_equal(self.reg_vtx, o.reg_vtx):
            return False, "reg_vtx"
        if not np.array_equal(self.reg_face, o.reg_face):
            return False, "reg_face"
        if len(self.lists) != len(o.lists):
            return False, "nlists"
        for l, (a, b) in enumerate(zip(self.lists, o.lists)):
            for name in ("type", "aux", "symbols", "hist", "type_hist"):
                x, y = getattr(a, name), getattr(b, name)
                if x.shape != y.shape or not np.array_equal(x, y):
                    return False, f"list {l} {name} {x.shape} vs {y.shape}"
        return True, ""


def streams_to_py(sp, copy: bool = True) -> StreamsPy:
    """hb_streams -> numpy.  copy=False returns views of the library's (page-locked) buffers: valid
    until hb_streams_free is called on `sp` -- what a C++ caller of the ABI sees, no extra pass."""
    return streams_struct_to_py(sp.contents, copy)


def streams_struct_to_py(s, copy: bool = True) -> StreamsPy:

    def arr(ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        if not ptr:   # an all-zero stream is returned as NULL (harry_b200.h, hb_attr_encode): a zero-stride view
            return np.broadcast_to(np.zeros(1, dtype=dtype), (n,))
        addr = C.cast(ptr, C.c_void_p).value
        buf = (C.c_uint8 * (n * np.dtype(dtype).itemsize)).from_address(addr)
        a = np.frombuffer(buf, dtype=dtype)
        return a.copy() if copy else a

    copied = 0   # bytes the library actually copied device -> host (NULL streams are not)

    def nb(ptr, n, itemsize):
        return n * itemsize if ptr and n else 0

    lists = []
    for l in range(s.nlists):
        ls = s.lists[l]
        copied += nb(ls.type, ls.n_emit, 1) + nb(ls.aux, ls.n_emit, 4) + nb(ls.symbols, ls.n_data * ls.sym_stride, 1) + (ls.sym_stride * 256 + 4) * 8
        lists.append(ListStreamsPy(
            type=arr(ls.type, ls.n_emit, np.uint8),
            aux=arr(ls.aux, ls.n_emit, np.uint32),
            symbols=arr(ls.symbols, ls.n_data * ls.sym_stride, np.uint8).reshape(ls.n_data, ls.sym_stride),
            hist=arr(ls.hist, ls.sym_stride * 256, np.uint64).reshape(ls.sym_stride, 256),
            type_hist=np.array([ls.type_hist[k] for k in range(4)], dtype=np.uint64)))
    copied += nb(s.reg_vtx, s.n_vtx, 2) + nb(s.reg_face, s.n_face, 2)
    out = StreamsPy(arr(s.reg_vtx, s.n_vtx, np.uint16), arr(s.reg_face, s.n_face, np.uint16), lists)
    out.nbytes_copied = copied
    return out


def residual_rows_from_streams(mesh: MeshArrays, st: StreamsPy) -> list:
    """Decode-side input: rows[k] of list l = k-th DATA residual row, scattered to the component
    offsets in storage type (what AttrDecoder reads into the row before decodeDelta,
    formats/hry/attrcode.h:459-461)."""
    out = []
    for l, la in enumerate(mesh.lists):
        lb = la.copy()
        lb.rows[:] = 0
        sym = st.lists[l].symbols
        pos = 0
        for j in range(la.ncomp):
            sz = TYPE_SIZE[la.stype(j)]
            off = la.offsets[j]
            lb.rows[: sym.shape[0], off:off + sz] = sym[:, pos:pos + sz]
            pos += sz
        out.append(lb)
    return out


def residual_rows_encoder_side(mesh: MeshArrays, st: StreamsPy) -> list:
    """Like residual_rows_from_streams, but for a mesh in ENCODER numbering: the k-th DATA residual
    of a list goes to the attribute row its emitting element is bound to (vertex and face lists;
    used by smoke() and tests to decode on the mesh that was just encoded)."""
    face_of = np.repeat(np.arange(mesh.nf, dtype=np.int64), np.diff(mesh.face_off.astype(np.int64)))
    out = []
    for l, la in enumerate(mesh.lists):
        lb = la.copy()
        lb.rows[:] = 0
        if la.target == T_VTX:
            h = mesh.face_off[mesh.order[:, 0]].astype(np.int64) + mesh.order[:, 1]
            ent = mesh.edges[h, 0].astype(np.int64)
            regs, off, lists, bind, nb = mesh.vtx_regs, mesh.off_reg_vtx, mesh.reg_vtxlist, mesh.bind_vtx, mesh.nb_vtx
        elif la.target == T_FACE:
            of = mesh.order_f if mesh.order_f is not None else np.stack([np.arange(mesh.nf), np.zeros(mesh.nf)], 1)
            ent = of[:, 0].astype(np.int64)
            regs, off, lists, bind, nb = mesh.face_regs, mesh.off_reg_face, mesh.reg_facelist, mesh.bind_face, mesh.nb_face
        else:
            raise NotImplementedError("corner lists: use the decoder-side mesh")
        slot = np.full(len(off) - 1, -1, dtype=np.int64)
        for r in range(len(off) - 1):
            for a in range(off[r + 1] - off[r]):
                if lists[off[r] + a] == l:
                    slot[r] = a
        a_of = slot[regs[ent]]
        ent = ent[a_of >= 0]
        rows_idx = bind[ent * nb + a_of[a_of >= 0]].astype(np.int64)
        _, first = np.unique(rows_idx, return_index=True)
        data_rows = rows_idx[np.sort(first)]          # row of the k-th DATA emission
        sym = st.lists[l].symbols
        assert sym.shape[0] == data_rows.shape[0]
        pos = 0
        for j in range(la.ncomp):
            sz = TYPE_SIZE[la.stype(j)]
            o = la.offsets[j]
            lb.rows[data_rows, o:o + sz] = sym[:, pos:pos + sz]
            pos += sz
        out.append(lb)
    return out


# ----------------------------------------------------------------------------------------------
# library loading
# ----------------------------------------------------------------------------------------------
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libharry_b200.so")
_lib = None


class HarryError(RuntimeError):
    pass


def load_library():
    """Load the CUDA library.  Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HarryError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(there is no CPU fallback for the attribute path)")
    lib = C.CDLL(LIB_PATH)
    vp, u32 = C.c_void_p, C.c_uint32
    lib.hb_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hb_ctx_create.restype = C.c_int
    lib.hb_ctx_create_prio.argtypes = [C.c_int, C.c_int, C.POINTER(vp)]
    lib.hb_ctx_create_prio.restype = C.c_int
    lib.hb_ctx_destroy.argtypes = [vp]
    lib.hb_ctx_destroy.restype = None
    lib.hb_last_error.argtypes = [vp]
    lib.hb_last_error.restype = C.c_char_p
    lib.hb_last_timing.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.hb_last_timing.restype = None
    lib.hb_ctx_set_row_cache.argtypes = [vp, C.c_int]
    lib.hb_ctx_set_row_cache.restype = C.c_int
    lib.hb_kernel_launches.argtypes = [vp]
    lib.hb_kernel_launches.restype = C.c_uint64
    lib.hb_h2d_bytes.argtypes = [vp]
    lib.hb_h2d_bytes.restype = C.c_uint64
    lib.hb_ctx_profile.argtypes = [vp, C.c_int]
    lib.hb_ctx_profile.restype = C.c_int
    lib.hb_ctx_profile_report.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.hb_ctx_profile_report.restype = C.c_int
    lib.hb_ctx_mark.argtypes = [vp, C.c_int]
    lib.hb_ctx_mark.restype = C.c_int
    lib.hb_ctx_elapsed.argtypes = [vp, C.c_int, C.c_int, C.POINTER(C.c_float)]
    lib.hb_ctx_elapsed.restype = C.c_int
    lib.hb_bounds.argtypes = [vp, C.POINTER(ListDesc), vp, vp]
    lib.hb_bounds.restype = C.c_int
    lib.hb_requant.argtypes = [vp, C.POINTER(ListDesc), C.POINTER(C.c_uint8), vp, vp]
    lib.hb_requant.restype = C.c_int
    lib.hb_twin_match.argtypes = [vp, u32, u32, vp, vp, u32, vp]
    lib.hb_twin_match.restype = C.c_int
    lib.hb_attr_encode.argtypes = [vp, C.POINTER(MeshDesc), C.POINTER(C.POINTER(Streams))]
    lib.hb_attr_encode.restype = C.c_int
    lib.hb_streams_free.argtypes = [C.POINTER(Streams)]
    lib.hb_streams_free.restype = None
    lib.hb_attr_decode.argtypes = [vp, C.POINTER(MeshDesc)]
    lib.hb_attr_decode.restype = C.c_int
    lib.hb_dmesh_upload.argtypes = [vp, C.POINTER(MeshDesc), C.POINTER(vp)]
    lib.hb_dmesh_upload.restype = C.c_int
    lib.hb_dmesh_upload_batch.argtypes = [vp, C.POINTER(MeshDesc), u32, C.POINTER(vp)]
    lib.hb_dmesh_upload_batch.restype = C.c_int
    lib.hb_dmesh_segments.argtypes = [vp]
    lib.hb_dmesh_segments.restype = u32
    lib.hb_dmesh_fetch_streams_batch.argtypes = [vp, C.POINTER(C.POINTER(BatchStreams))]
    lib.hb_dmesh_fetch_streams_batch.restype = C.c_int
    lib.hb_dmesh_fetch_rows_seg.argtypes = [vp, u32, u32, vp]
    lib.hb_dmesh_fetch_rows_seg.restype = C.c_int
    lib.hb_batch_streams_free.argtypes = [C.POINTER(BatchStreams)]
    lib.hb_batch_streams_free.restype = None
    lib.hb_encode_batch.argtypes = [vp, C.POINTER(MeshDesc), u32, C.POINTER(QuantReq), u32, C.POINTER(vp), C.POINTER(C.POINTER(BatchStreams))]
    lib.hb_encode_batch.restype = C.c_int
    lib.hb_decode_batch.argtypes = [vp, C.POINTER(MeshDesc), u32, C.POINTER(DequantReq), u32]
    lib.hb_decode_batch.restype = C.c_int
    lib.hb_dmesh_free.argtypes = [vp]
    lib.hb_dmesh_free.restype = None
    lib.hb_dmesh_quantize.argtypes = [vp, u32, C.POINTER(C.c_uint8), C.POINTER(C.c_uint8)]
    lib.hb_dmesh_quantize.restype = C.c_int
    lib.hb_dmesh_dequantize.argtypes = [vp, u32]
    lib.hb_dmesh_dequantize.restype = C.c_int
    lib.hb_dmesh_encode.argtypes = [vp]
    lib.hb_dmesh_encode.restype = C.c_int
    lib.hb_dmesh_fetch_streams.argtypes = [vp, C.POINTER(C.POINTER(Streams))]
    lib.hb_dmesh_fetch_streams.restype = C.c_int
    lib.hb_dmesh_set_bounds.argtypes = [vp, u32, vp, vp, vp]
    lib.hb_dmesh_set_bounds.restype = C.c_int
    lib.hb_dmesh_snapshot.argtypes = [vp]
    lib.hb_dmesh_snapshot.restype = C.c_int
    lib.hb_dmesh_restore.argtypes = [vp]
    lib.hb_dmesh_restore.restype = C.c_int
    lib.hb_dmesh_decode.argtypes = [vp]
    lib.hb_dmesh_decode.restype = C.c_int
    lib.hb_dmesh_fetch_rows.argtypes = [vp, u32, vp]
    lib.hb_dmesh_fetch_rows.restype = C.c_int
    lib.hb_dmesh_decode_stats.argtypes = [vp, u32, C.POINTER(C.c_uint64)]
    lib.hb_dmesh_decode_stats.restype = C.c_int
    lib.hb_dmesh_fetch_bounds.argtypes = [vp, u32, vp, vp, vp]
    lib.hb_dmesh_fetch_bounds.restype = C.c_int
    lib.hb_ctx_wait.argtypes = [vp, vp]
    lib.hb_ctx_wait.restype = C.c_int
    lib.hb_ctx_sync.argtypes = [vp]
    lib.hb_ctx_sync.restype = C.c_int
    _lib = lib
    return lib


EXPORTED_SYMBOLS = [
    "hb_ctx_create", "hb_ctx_create_prio", "hb_ctx_destroy", "hb_last_error", "hb_last_timing", "hb_kernel_launches", "hb_h2d_bytes",
    "hb_ctx_profile", "hb_ctx_profile_report", "hb_ctx_mark", "hb_ctx_elapsed",
    "hb_bounds", "hb_requant", "hb_attr_encode", "hb_streams_free", "hb_attr_decode", "hb_twin_match",
    "hb_dmesh_upload", "hb_dmesh_free", "hb_dmesh_quantize", "hb_dmesh_dequantize", "hb_dmesh_encode",
    "hb_dmesh_fetch_streams", "hb_dmesh_set_bounds", "hb_dmesh_snapshot", "hb_dmesh_restore", "hb_dmesh_decode", "hb_dmesh_fetch_rows",
    "hb_dmesh_fetch_bounds", "hb_dmesh_decode_stats", "hb_ctx_sync",
    "hb_dmesh_upload_batch", "hb_dmesh_segments", "hb_dmesh_fetch_streams_batch", "hb_dmesh_fetch_rows_seg",
    "hb_encode_batch", "hb_decode_batch", "hb_batch_streams_free", "hb_ctx_set_row_cache", "hb_ctx_wait",
]


def _desc_array(meshes):
    arr = (MeshDesc * len(meshes))()
    for i, m in enumerate(meshes):
        arr[i] = m.to_desc()
    return arr


class Context:
    """RAII wrapper of hb_ctx with numpy-level calls mirroring the reference operators."""

    def __init__(self, device: int = 0, high_priority: bool = False):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.hb_ctx_create_prio(device, 1 if high_priority else 0, C.byref(h))
        if rc != 0:
            raise HarryError(f"hb_ctx_create failed ({rc}): {self.lib.hb_last_error(None).decode()}")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.hb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise HarryError(f"{what} failed ({rc}): {self.lib.hb_last_error(self.h).decode()}")

    def timing(self):
        k, c = C.c_float(), C.c_float()
        self.lib.hb_last_timing(self.h, C.byref(k), C.byref(c))
        return k.value, c.value

    def set_row_cache(self, enable: bool):
        """Keep the device copy of a list's rows between the calls of one pipeline (harry_b200.h, hb_ctx_set_row_cache)."""
        self._check(self.lib.hb_ctx_set_row_cache(self.h, 1 if enable else 0), "hb_ctx_set_row_cache")

    def launches(self) -> int:
        return int(self.lib.hb_kernel_launches(self.h))

    def h2d_bytes(self) -> int:
        """Bytes of mesh arrays and rows copied host -> device since the context was created."""
        return int(self.lib.hb_h2d_bytes(self.h))

    def sync(self):
        self._check(self.lib.hb_ctx_sync(self.h), "hb_ctx_sync")

    def wait(self, other: "Context"):
        """what is queued on `other` so far happens before what is queued on this context from now on"""
        self._check(self.lib.hb_ctx_wait(self.h, other.h), "hb_ctx_wait")

    def profile(self, enable: bool):
        self._check(self.lib.hb_ctx_profile(self.h, 1 if enable else 0), "hb_ctx_profile")

    def profile_report(self) -> dict:
        """{kernel name: (launches, total ms)} since profiling was enabled / last report."""
        buf = C.create_string_buffer(1 << 16)
        self._check(self.lib.hb_ctx_profile_report(self.h, buf, len(buf)), "hb_ctx_profile_report")
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def mark(self, idx: int):
        self._check(self.lib.hb_ctx_mark(self.h, idx), "hb_ctx_mark")

    def elapsed(self, a: int, b: int) -> float:
        ms = C.c_float()
        self._check(self.lib.hb_ctx_elapsed(self.h, a, b, C.byref(ms)), "hb_ctx_elapsed")
        return ms.value

    # quant::set_bounds
    def bounds(self, la: ListArrays):
        d = la.to_desc()
        mn = np.zeros(max(1, la.stride), dtype=np.uint8)
        mx = np.zeros(max(1, la.stride), dtype=np.uint8)
        self._check(self.lib.hb_bounds(self.h, C.byref(d), mn.ctypes.data, mx.ctypes.data), "hb_bounds")
        return mn[: la.stride], mx[: la.stride]

    # quant::requant(Attr&, Fmt)
    def requant(self, la: ListArrays, new_quant, min_row, scale_row):
        d = la.to_desc()
        nq = (C.c_uint8 * HB_MAX_COMP)(*list(new_quant))
        mn = np.ascontiguousarray(min_row, dtype=np.uint8)
        sc = np.ascontiguousarray(scale_row, dtype=np.uint8)
        self._check(self.lib.hb_requant(self.h, C.byref(d), nq, mn.ctypes.data, sc.ctypes.data), "hb_requant")
        la.sync_from_desc(d)

    # mesh::Builder::add_edge over all faces (structs/conn.h:164-214)
    def twin_match(self, nv: int, face_off: np.ndarray, org: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """org: (ne,) uint32 origins, or (ne, 3) uint32 edge records whose column 0 holds them (matched in place when
        `out` is the same array).  Returns (ne, 3) uint32 records {org, twin_face, twin_edge}."""
        face_off = np.ascontiguousarray(face_off, dtype=np.uint32)
        nf = int(face_off.shape[0]) - 1
        ne = int(face_off[nf]) if nf > 0 else 0
        if org.dtype != np.uint32 or not org.flags.c_contiguous:
            org = np.ascontiguousarray(org, dtype=np.uint32)
        stride = 12 if org.ndim == 2 else 4
        if org.shape[0] != ne or (org.ndim == 2 and org.shape[1] != 3):
            raise HarryError(f"twin_match: org has shape {org.shape}, expected ({ne},) or ({ne}, 3)")
        if out is None:
            out = np.zeros((ne, 3), dtype=np.uint32)
        self._check(self.lib.hb_twin_match(self.h, nv, max(nf, 0), face_off.ctypes.data, org.ctypes.data, stride, out.ctypes.data),
                    "hb_twin_match")
        return out

    # AttrCoder::encode
    def attr_encode(self, mesh: MeshArrays) -> StreamsPy:
        d = mesh.to_desc()
        sp = C.POINTER(Streams)()
        self._check(self.lib.hb_attr_encode(self.h, C.byref(d), C.byref(sp)), "hb_attr_encode")
        try:
            return streams_to_py(sp)
        finally:
            self.lib.hb_streams_free(sp)

    def attr_encode_view(self, mesh: MeshArrays):
        """Same call, zero-copy: returns (streams viewing the library's buffers, release callable)."""
        d = mesh.to_desc()
        sp = C.POINTER(Streams)()
        self._check(self.lib.hb_attr_encode(self.h, C.byref(d), C.byref(sp)), "hb_attr_encode")
        return streams_to_py(sp, copy=False), (lambda: self.lib.hb_streams_free(sp))

    # AttrDecoder::decode (value reconstruction)
    def attr_decode(self, mesh: MeshArrays) -> None:
        d = mesh.to_desc()
        self._check(self.lib.hb_attr_decode(self.h, C.byref(d)), "hb_attr_decode")


    # ---- batches of independent meshes (host buffers, pipelined in groups) ----------------------
    def encode_batch(self, meshes, quant=(), want_bounds: bool = True, copy: bool = True):
        """hb_encode_batch: per mesh set_bounds + set_scale + requant of the lists in `quant`
        [(list, new_quant, groups)], then AttrCoder::encode.  Returns ([StreamsPy per mesh],
        [bounds array (n, 3, stride) uint8 per request], release callable when copy=False)."""
        n = len(meshes)
        descs = _desc_array(meshes)
        nq = len(quant)
        reqs = (QuantReq * max(1, nq))()
        bounds, bptr = [], (C.c_void_p * max(1, nq))()
        for k, (l, new_quant, groups) in enumerate(quant):
            reqs[k].list = l
            for j, v in enumerate(new_quant):
                reqs[k].new_quant[j] = v
            for j, v in enumerate(groups):
                reqs[k].groups[j] = v
            b = np.zeros((n, 3, max(1, meshes[0].lists[l].stride)), dtype=np.uint8)
            bounds.append(b)
            bptr[k] = b.ctypes.data if want_bounds else None
        bp = C.POINTER(BatchStreams)()
        self._check(self.lib.hb_encode_batch(self.h, descs, n, reqs, nq, bptr, C.byref(bp)), "hb_encode_batch")
        out = [streams_struct_to_py(bp.contents.mesh[i], copy) for i in range(n)]
        if copy:
            self.lib.hb_batch_streams_free(bp)
            return out, bounds
        return out, bounds, (lambda: self.lib.hb_batch_streams_free(bp))

    def decode_batch(self, meshes, dequant=()):
        """hb_decode_batch: per mesh AttrDecoder::decode, then requant(clear) of the lists in `dequant`
        [(list, bounds array (n, 3, stride) uint8)].  Rows of every list are updated in place."""
        n = len(meshes)
        descs = _desc_array(meshes)
        nq = len(dequant)
        reqs = (DequantReq * max(1, nq))()
        keep = []
        for k, (l, b) in enumerate(dequant):
            b = np.ascontiguousarray(b, dtype=np.uint8)
            keep.append(b)
            reqs[k].list = l
            reqs[k].bounds = b.ctypes.data
        self._check(self.lib.hb_decode_batch(self.h, descs, n, reqs, nq), "hb_decode_batch")
        for m in meshes:
            for l, _ in dequant:
                m.lists[l].quants = [0] * m.lists[l].ncomp


class DeviceMesh:
    """Device-resident mesh (hb_dmesh): upload once, run stages as kernels only.  With a list of
    meshes of one schema: a batch -- ONE device mesh whose stages run once over all of them."""

    def __init__(self, ctx: Context, mesh):
        self.ctx = ctx
        self.meshes = list(mesh) if isinstance(mesh, (list, tuple)) else [mesh]
        self.mesh = self.meshes[0]
        descs = _desc_array(self.meshes)
        h = C.c_void_p()
        ctx._check(ctx.lib.hb_dmesh_upload_batch(ctx.h, descs, len(self.meshes), C.byref(h)), "hb_dmesh_upload_batch")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.hb_dmesh_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def quantize(self, l: int, new_quant, groups):
        nq = (C.c_uint8 * HB_MAX_COMP)(*list(new_quant))
        gr = (C.c_uint8 * HB_MAX_COMP)(*list(groups))
        self.ctx._check(self.ctx.lib.hb_dmesh_quantize(self.h, l, nq, gr), "hb_dmesh_quantize")

    def dequantize(self, l: int):
        self.ctx._check(self.ctx.lib.hb_dmesh_dequantize(self.h, l), "hb_dmesh_dequantize")

    def encode(self):
        self.ctx._check(self.ctx.lib.hb_dmesh_encode(self.h), "hb_dmesh_encode")

    def fetch_streams(self) -> StreamsPy:
        sp = C.POINTER(Streams)()
        self.ctx._check(self.ctx.lib.hb_dmesh_fetch_streams(self.h, C.byref(sp)), "hb_dmesh_fetch_streams")
        try:
            return streams_to_py(sp)
        finally:
            self.ctx.lib.hb_streams_free(sp)

    def fetch_streams_batch(self) -> list:
        bp = C.POINTER(BatchStreams)()
        self.ctx._check(self.ctx.lib.hb_dmesh_fetch_streams_batch(self.h, C.byref(bp)), "hb_dmesh_fetch_streams_batch")
        try:
            return [streams_struct_to_py(bp.contents.mesh[i]) for i in range(bp.contents.n)]
        finally:
            self.ctx.lib.hb_batch_streams_free(bp)

    def set_bounds(self, l: int, mn=None, mx=None, sc=None):
        """Rows of all meshes back to back for a batch: arrays of shape (n, stride)."""
        keep = [None if a is None else np.ascontiguousarray(a, dtype=np.uint8) for a in (mn, mx, sc)]
        ptr = lambda a: None if a is None else a.ctypes.data
        self.ctx._check(self.ctx.lib.hb_dmesh_set_bounds(self.h, l, ptr(keep[0]), ptr(keep[1]), ptr(keep[2])), "hb_dmesh_set_bounds")
        del keep

    def snapshot(self):
        self.ctx._check(self.ctx.lib.hb_dmesh_snapshot(self.h), "hb_dmesh_snapshot")

    def restore(self):
        self.ctx._check(self.ctx.lib.hb_dmesh_restore(self.h), "hb_dmesh_restore")

    def decode(self):
        self.ctx._check(self.ctx.lib.hb_dmesh_decode(self.h), "hb_dmesh_decode")

    def fetch_rows(self, l: int, seg: int = 0) -> np.ndarray:
        la = self.meshes[seg].lists[l]
        out = np.zeros((la.nrows, la.stride), dtype=np.uint8)
        self.ctx._check(self.ctx.lib.hb_dmesh_fetch_rows_seg(self.h, seg, l, out.ctypes.data), "hb_dmesh_fetch_rows_seg")
        return out

    def decode_stats(self, l: int) -> list:
        out = (C.c_uint64 * 8)()
        self.ctx._check(self.ctx.lib.hb_dmesh_decode_stats(self.h, l, out), "hb_dmesh_decode_stats")
        return [int(v) for v in out]

    def fetch_bounds(self, l: int):
        """(min, max, scale) rows; for a batch arrays of shape (n, stride)."""
        la = self.mesh.lists[l]
        n, ns = max(1, la.stride), len(self.meshes)
        mn, mx, sc = (np.zeros((ns, n), dtype=np.uint8) for _ in range(3))
        self.ctx._check(self.ctx.lib.hb_dmesh_fetch_bounds(self.h, l, mn.ctypes.data, mx.ctypes.data, sc.ctypes.data),
                        "hb_dmesh_fetch_bounds")
        if ns == 1:
            return mn[0, : la.stride], mx[0, : la.stride], sc[0, : la.stride]
        return mn[:, : la.stride], mx[:, : la.stride], sc[:, : la.stride]
