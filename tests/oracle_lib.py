"""ctypes bindings of the TEST ORACLE: oracle/libharry_oracle.so (C restatement) and
oracle/_ref/libharry_ref.so (the unmodified reference behind a harness).  Only tests/, smoke()
and bench.py's CPU-baseline legs import this module; the product package never does."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from harry_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libharry_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libharry_ref.so")
REF_CLI = os.path.join(ORACLE_DIR, "_ref", "harry")

_oracle = None
_ref = None


class quiet_stdout:
    """The reference prints progress bars and statistics on stdout (utils/progress.h:102-136,
    formats/obj/reader.cc); keep them away from our stdout (bench.py prints one JSON line there)."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)
        return self

    def __exit__(self, *exc):
        os.dup2(self.saved, 1)
        os.close(self.saved)
        os.close(self.null)
        return False


def build_oracle():
    """Compile the C restatement (gcc only, no reference needed)."""
    subprocess.run(["make", "-C", ORACLE_DIR, "libharry_oracle.so"], check=True, capture_output=True)


def oracle():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build_oracle()
        lib = C.CDLL(ORACLE_SO)
        vp = C.c_void_p
        lib.ho_last_error.restype = C.c_char_p
        lib.ho_bounds.argtypes = [C.POINTER(capi.ListDesc), vp, vp]
        lib.ho_scale.argtypes = [C.POINTER(capi.ListDesc), C.POINTER(C.c_uint8), vp, vp, vp]
        lib.ho_requant.argtypes = [C.POINTER(capi.ListDesc), C.POINTER(C.c_uint8), vp, vp]
        lib.ho_attr_encode.argtypes = [C.POINTER(capi.MeshDesc), C.POINTER(C.POINTER(capi.Streams))]
        lib.ho_streams_free.argtypes = [C.POINTER(capi.Streams)]
        lib.ho_streams_free.restype = None
        lib.ho_attr_decode.argtypes = [C.POINTER(capi.MeshDesc)]
        for f in ("ho_predict", "ho_encode_delta", "ho_decode_delta"):
            getattr(lib, f).restype = C.c_uint64
        lib.ho_predict.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        lib.ho_encode_delta.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        lib.ho_decode_delta.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_int]
        lib.ho_twin_match.argtypes = [C.c_uint32, C.c_uint32, vp, vp, C.c_uint32, vp]
        lib.ho_msb.argtypes = [C.c_uint32]
        lib.ho_msb.restype = C.c_uint32
        _oracle = lib
    return _oracle


def _ocheck(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed ({rc}): {oracle().ho_last_error().decode()}")


def o_bounds(la: capi.ListArrays):
    d = la.to_desc()
    n = max(1, la.stride)
    mn, mx = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    _ocheck(oracle().ho_bounds(C.byref(d), mn.ctypes.data, mx.ctypes.data), "bounds")
    return mn[: la.stride], mx[: la.stride]


def o_scale(la: capi.ListArrays, mn, mx):
    d = la.to_desc()
    g = (C.c_uint8 * capi.HB_MAX_COMP)(*list(la.groups))
    sc = np.zeros(max(1, la.stride), np.uint8)
    mn = np.ascontiguousarray(mn)
    mx = np.ascontiguousarray(mx)
    _ocheck(oracle().ho_scale(C.byref(d), g, mn.ctypes.data, mx.ctypes.data, sc.ctypes.data), "scale")
    return sc[: la.stride]


def o_requant(la: capi.ListArrays, new_quant, mn, sc):
    d = la.to_desc()
    nq = (C.c_uint8 * capi.HB_MAX_COMP)(*list(new_quant))
    mn = np.ascontiguousarray(mn)
    sc = np.ascontiguousarray(sc)
    _ocheck(oracle().ho_requant(C.byref(d), nq, mn.ctypes.data, sc.ctypes.data), "requant")
    la.sync_from_desc(d)


def o_twin_match(nv: int, face_off: np.ndarray, org: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    """Same contract as capi.Context.twin_match."""
    face_off = np.ascontiguousarray(face_off, dtype=np.uint32)
    nf = int(face_off.shape[0]) - 1
    ne = int(face_off[nf]) if nf > 0 else 0
    org = np.ascontiguousarray(org, dtype=np.uint32)
    if out is None:
        out = np.zeros((ne, 3), dtype=np.uint32)
    _ocheck(oracle().ho_twin_match(nv, max(nf, 0), face_off.ctypes.data, org.ctypes.data, 12 if org.ndim == 2 else 4, out.ctypes.data),
            "twin_match")
    return out


def o_attr_encode(mesh: capi.MeshArrays) -> capi.StreamsPy:
    d = mesh.to_desc()
    sp = C.POINTER(capi.Streams)()
    _ocheck(oracle().ho_attr_encode(C.byref(d), C.byref(sp)), "attr_encode")
    try:
        return capi.streams_to_py(sp)
    finally:
        oracle().ho_streams_free(sp)


def o_attr_decode(mesh: capi.MeshArrays) -> None:
    d = mesh.to_desc()
    _ocheck(oracle().ho_attr_decode(C.byref(d)), "attr_decode")


# ----------------------------------------------------------------------------------------------
# the real reference behind oracle/ref_harness.cc
# ----------------------------------------------------------------------------------------------
def have_ref() -> bool:
    return os.path.exists(REF_SO)


def ref():
    global _ref
    if _ref is None:
        lib = C.CDLL(REF_SO)
        vp = C.c_void_p
        lib.ref_last_error.restype = C.c_char_p
        lib.ref_read.argtypes = [C.c_char_p]
        lib.ref_read.restype = vp
        lib.ref_free.argtypes = [vp]
        lib.ref_free.restype = None
        lib.ref_write.argtypes = [vp, C.c_char_p]
        lib.ref_set_bounds.argtypes = [vp]
        lib.ref_requant.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.c_int]
        lib.ref_traverse.argtypes = [vp]
        lib.ref_desc.argtypes = [vp, C.POINTER(capi.MeshDesc)]
        lib.ref_groups.argtypes = [vp, C.c_int]
        lib.ref_groups.restype = C.POINTER(C.c_uint8)
        lib.ref_bounds_row.argtypes = [vp, C.c_int, C.c_int]
        lib.ref_bounds_row.restype = vp
        lib.ref_set_scale.argtypes = [vp, C.c_int]
        lib.ref_has_logged.argtypes = [vp]
        lib.ref_attr_encode.argtypes = [vp, C.POINTER(C.POINTER(capi.Streams))]
        lib.ref_logged_streams.argtypes = [vp, C.POINTER(C.POINTER(capi.Streams))]
        lib.ref_streams_free.argtypes = [C.POINTER(capi.Streams)]
        lib.ref_streams_free.restype = None
        lib.ref_time_path.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
        lib.ref_time_decode.argtypes = [C.c_char_p, C.POINTER(C.c_double)]
        if hasattr(lib, "ref_snapshot"):     # several timed steps on one prepared mesh
            lib.ref_snapshot.argtypes = [vp]
            lib.ref_restore.argtypes = [vp]
            lib.ref_time_encode_step.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_double)]
            lib.ref_decode_open.argtypes = [C.c_char_p]
            lib.ref_decode_open.restype = vp
            lib.ref_decode_close.argtypes = [vp]
            lib.ref_decode_close.restype = None
            lib.ref_decode_step.argtypes = [vp, C.POINTER(C.c_double)]
        if hasattr(lib, "ref_twin_match"):   # harness builds older than the twin matching row lack it
            lib.ref_twin_match.argtypes = [C.c_uint32, vp, vp, vp, C.POINTER(C.c_double)]
        _ref = lib
    return _ref


class RefMesh:
    """A mesh::Mesh living inside the reference harness."""

    def __init__(self, path: str):
        self.lib = ref()
        with quiet_stdout():
            self.h = self.lib.ref_read(path.encode())
        if not self.h:
            raise RuntimeError(f"ref_read({path}): {self.lib.ref_last_error().decode()}")

    def close(self):
        if self.h:
            self.lib.ref_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"reference {what}: {self.lib.ref_last_error().decode()}")

    def requant(self, loq, clear=False):
        flat = [int(x) for t in loq for x in t]
        arr = (C.c_int * max(1, len(flat)))(*flat)
        self._check(self.lib.ref_requant(self.h, len(loq), arr, 1 if clear else 0), "requant")

    def set_bounds(self):
        self._check(self.lib.ref_set_bounds(self.h), "set_bounds")

    def traverse(self):
        with quiet_stdout():
            rc = self.lib.ref_traverse(self.h)
        self._check(rc, "traverse")

    def write(self, path: str):
        with quiet_stdout():
            rc = self.lib.ref_write(self.h, path.encode())
        self._check(rc, "write")

    def arrays(self) -> capi.MeshArrays:
        """Deep numpy copy of the flattened mesh (current state)."""
        d = capi.MeshDesc()
        self._check(self.lib.ref_desc(self.h, C.byref(d)), "desc")

        def groups(l):
            n = d.lists[l].ncomp
            p = self.lib.ref_groups(self.h, l)
            return [int(p[j]) for j in range(n)]

        return capi.mesh_from_desc(d, groups)

    def bounds_row(self, l: int, which: int, stride: int) -> np.ndarray:
        p = self.lib.ref_bounds_row(self.h, l, which)
        if stride == 0:
            return np.zeros(0, np.uint8)
        return np.frombuffer((C.c_uint8 * stride).from_address(p), dtype=np.uint8).copy()

    def set_scale(self, l: int):
        self.lib.ref_set_scale(self.h, l)

    def attr_encode(self) -> capi.StreamsPy:
        sp = C.POINTER(capi.Streams)()
        self._check(self.lib.ref_attr_encode(self.h, C.byref(sp)), "attr_encode")
        try:
            return capi.streams_to_py(sp)
        finally:
            self.lib.ref_streams_free(sp)

    def logged_streams(self) -> capi.StreamsPy:
        sp = C.POINTER(capi.Streams)()
        self._check(self.lib.ref_logged_streams(self.h, C.byref(sp)), "logged_streams")
        try:
            return capi.streams_to_py(sp)
        finally:
            self.lib.ref_streams_free(sp)

    def time_path(self, loq):
        flat = [int(x) for t in loq for x in t]
        arr = (C.c_int * max(1, len(flat)))(*flat)
        t = (C.c_double * 6)()
        with quiet_stdout():
            rc = self.lib.ref_time_path(self.h, len(loq), arr, t)
        self._check(rc, "time_path")
        return list(t)


    def snapshot(self):
        self._check(self.lib.ref_snapshot(self.h), "snapshot")

    def restore(self):
        self._check(self.lib.ref_restore(self.h), "restore")

    def time_encode_step(self, loq):
        """[set_bounds, requant, 0, AttrCoder<NullWriter>::encode, 0, 0] seconds; works in place (restore() first)."""
        flat = [int(x) for t in loq for x in t]
        arr = (C.c_int * max(1, len(flat)))(*flat)
        t = (C.c_double * 6)()
        with quiet_stdout():
            rc = self.lib.ref_time_encode_step(self.h, len(loq), arr, t)
        self._check(rc, "time_encode_step")
        return list(t)


class RefDecoder:
    """A .hry read once by the reference's decoder; step() times AttrDecoder<Replay>::decode + requant(clear)."""

    def __init__(self, hry_path: str):
        self.lib = ref()
        with quiet_stdout():
            self.h = self.lib.ref_decode_open(hry_path.encode())
        if not self.h:
            raise RuntimeError(f"ref_decode_open({hry_path}): {self.lib.ref_last_error().decode()}")

    def step(self):
        t = (C.c_double * 6)()
        with quiet_stdout():
            rc = self.lib.ref_decode_step(self.h, t)
        if rc != 0:
            raise RuntimeError(f"reference decode_step: {self.lib.ref_last_error().decode()}")
        return list(t)

    def close(self):
        if self.h:
            self.lib.ref_decode_close(self.h)
            self.h = None


def have_ref_twin() -> bool:
    return have_ref() and hasattr(ref(), "ref_twin_match")


def ref_twin_match(face_off: np.ndarray, org: np.ndarray):
    """The reference's conn::Builder (structs/conn.h:172-233) on a face list: ((ne, 3) uint32 records, seconds)."""
    face_off = np.ascontiguousarray(face_off, dtype=np.uint32)
    org = np.ascontiguousarray(org, dtype=np.uint32)
    nf = int(face_off.shape[0]) - 1
    out = np.zeros((int(face_off[nf]) if nf > 0 else 0, 3), dtype=np.uint32)
    sec = C.c_double()
    if ref().ref_twin_match(max(nf, 0), face_off.ctypes.data, org.ctypes.data, out.ctypes.data, C.byref(sec)) != 0:
        raise RuntimeError(f"reference twin_match: {ref().ref_last_error().decode()}")
    out[:, 2] &= 0xFFFF   # the pad bytes of fepair are not initialised by the reference
    return out, sec.value


def ref_time_decode(hry_path: str):
    t = (C.c_double * 6)()
    with quiet_stdout():
        rc = ref().ref_time_decode(hry_path.encode(), t)
    if rc != 0:
        raise RuntimeError(f"reference time_decode: {ref().ref_last_error().decode()}")
    return list(t)
