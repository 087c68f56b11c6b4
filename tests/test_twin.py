"""Twin matching (mesh::Builder, structs/conn.h:164-214).

CPU: the oracle restatement (ho_twin_match) against the committed golden twin tables the reference's own
reader built (tests/golden/twin/*.npz, tests/golden/*.npz) and, where oracle/_ref exists, against the
reference run live.  GPU: hb_twin_match through the C ABI against the same vectors and against the oracle,
bit for bit, plus size-independent properties at the full BASELINE size."""
import glob
import os

import numpy as np
import pytest

import golden_io
import oracle_lib as ol
from cases import CASES, CONFIG1
from harry_b200 import capi, meshgen

HERE = os.path.dirname(__file__)
TWIN_GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "twin", "*.npz")))
TWIN_IDS = [os.path.basename(p)[:-4] for p in TWIN_GOLDEN]
CASE_GOLDEN = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))
CASE_IDS = [os.path.basename(p)[:-4] for p in CASE_GOLDEN]
needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libharry_ref.so not built")

# generated inputs for oracle-vs-CUDA comparisons (no reference needed): name -> PolyMesh
GENERATED = {
    "sphere_40x80": lambda: meshgen.uv_sphere(40, 80),
    "poly_grid_24": lambda: meshgen.poly_grid(24),
    "irregular_48": lambda: meshgen.tri_irregular(48, 11),
    "cones_open": lambda: meshgen.cones(40, 100, seed=4, open_every=2),
    "soup_30": lambda: meshgen.soup(30, 400, 1, 1, 6),
    "soup_5": lambda: meshgen.soup(5, 3000, 4, 1, 6),
    "soup_2": lambda: meshgen.soup(2, 700, 9, 1, 4),
    "soup_1000": lambda: meshgen.soup(1000, 20000, 5, 3, 4),
    "soup_wide_faces": lambda: meshgen.soup(300, 40, 6, 200, 900),
}


def load_twin_golden(path):
    z = np.load(path)
    return int(z["nv"][0]), z["face_off"], z["edges"]


def check_table_properties(face_off, edges):
    """What holds for every table the Builder can produce: twin is an involution, a matched pair runs over the
    same vertex pair in opposite directions, borders point at themselves."""
    face_off = face_off.astype(np.int64)
    ne = int(face_off[-1])
    h = np.arange(ne, dtype=np.int64)
    f = np.searchsorted(face_off, h, side="right") - 1
    nxt = np.where(h + 1 == face_off[f + 1], face_off[f], h + 1)
    org = edges[:, 0].astype(np.int64)
    dst = org[nxt]
    t = face_off[edges[:, 1].astype(np.int64)] + edges[:, 2].astype(np.int64)
    assert np.all(edges[:, 2] < (face_off[edges[:, 1].astype(np.int64) + 1] - face_off[edges[:, 1].astype(np.int64)]))
    assert np.array_equal(t[t], h), "twin is not an involution"
    matched = t != h
    assert np.array_equal(org[matched], dst[t[matched]]) and np.array_equal(dst[matched], org[t[matched]])
    return int(np.count_nonzero(~matched))


# ---- CPU: oracle against the reference ---------------------------------------------------------
def test_twin_fixtures_present():
    assert len(TWIN_GOLDEN) >= 7


@pytest.mark.parametrize("path", TWIN_GOLDEN, ids=TWIN_IDS)
def test_oracle_twin_golden(path):
    nv, face_off, edges = load_twin_golden(path)
    got = ol.o_twin_match(nv, face_off, np.ascontiguousarray(edges[:, 0]))
    assert np.array_equal(got, edges)
    check_table_properties(face_off, got)


@pytest.mark.parametrize("path", CASE_GOLDEN, ids=CASE_IDS)
def test_oracle_twin_case_golden(path):
    raw = golden_io.GoldenCase(path).raw          # the mesh as the reference's reader left it
    got = ol.o_twin_match(raw.nv, raw.face_off, np.ascontiguousarray(raw.edges[:, 0]))
    assert np.array_equal(got, raw.edges)


def test_oracle_twin_in_place():
    nv, face_off, edges = load_twin_golden(TWIN_GOLDEN[0])
    rec = np.zeros_like(edges)
    rec[:, 0] = edges[:, 0]
    rec[:, 1:] = 0xdeadbeef                        # whatever the caller left in the twin words is ignored
    ol.o_twin_match(nv, face_off, rec, out=rec)
    assert np.array_equal(rec, edges)


def test_oracle_twin_rejects_bad_vertex():
    face_off = np.array([0, 3], np.uint32)
    with pytest.raises(RuntimeError):
        ol.o_twin_match(3, face_off, np.array([0, 1, 3], np.uint32))


def test_oracle_twin_face_size_limit():
    """a face may have 65535 corners (fepair::e is 16 bits wide), not more"""
    n = 65535
    org = np.arange(n, dtype=np.uint32)
    got = ol.o_twin_match(n, np.array([0, n], np.uint32), org)
    assert np.array_equal(got[:, 2], np.arange(n, dtype=np.uint32)) and np.all(got[:, 1] == 0)   # all borders
    with pytest.raises(RuntimeError):
        ol.o_twin_match(n + 1, np.array([0, n + 1], np.uint32), np.arange(n + 1, dtype=np.uint32))


@needs_ref
@pytest.mark.parametrize("name", list(CASES.keys()) + [CONFIG1[0]])
def test_oracle_twin_live_reference(workdir, name):
    gen = CONFIG1[1] if name == CONFIG1[0] else CASES[name][0]
    rm = ol.RefMesh(gen(workdir))
    m = rm.arrays()
    rm.close()
    assert np.array_equal(ol.o_twin_match(m.nv, m.face_off, np.ascontiguousarray(m.edges[:, 0])), m.edges)


@needs_ref
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_oracle_twin_live_reference_soup(workdir, seed):
    pm = meshgen.soup(4 + 3 * (seed % 5), 1500, seed, 1, 7)
    p = os.path.join(workdir, f"soup{seed}.ply")
    meshgen.write_ply(p, pm)
    rm = ol.RefMesh(p)
    m = rm.arrays()
    rm.close()
    assert np.array_equal(ol.o_twin_match(m.nv, m.face_off, np.ascontiguousarray(m.edges[:, 0])), m.edges)


@pytest.mark.skipif(not ol.have_ref_twin(), reason="oracle/_ref/libharry_ref.so lacks ref_twin_match")
@pytest.mark.parametrize("name", list(GENERATED.keys()))
def test_oracle_twin_vs_reference_builder(name):
    """the reference's conn::Builder driven directly (no file in between) on the generated inputs of the GPU tests"""
    pm = GENERATED[name]()
    want, _ = ol.ref_twin_match(pm.face_off, pm.face_idx)
    assert np.array_equal(ol.o_twin_match(pm.nv, pm.face_off, pm.face_idx), want)


@pytest.mark.skipif(not ol.have_ref_twin(), reason="oracle/_ref/libharry_ref.so lacks ref_twin_match")
def test_oracle_twin_vs_reference_builder_1m():
    pm = meshgen.uv_sphere(708, 1412)
    want, _ = ol.ref_twin_match(pm.face_off, pm.face_idx)
    assert np.array_equal(ol.o_twin_match(pm.nv, pm.face_off, pm.face_idx), want)


# ---- GPU: CUDA path through the C ABI ----------------------------------------------------------
@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.gpu
@pytest.mark.parametrize("path", TWIN_GOLDEN, ids=TWIN_IDS)
def test_cuda_twin_golden(ctx, path):
    nv, face_off, edges = load_twin_golden(path)
    assert np.array_equal(ctx.twin_match(nv, face_off, np.ascontiguousarray(edges[:, 0])), edges)


@pytest.mark.gpu
@pytest.mark.parametrize("path", CASE_GOLDEN, ids=CASE_IDS)
def test_cuda_twin_case_golden(ctx, path):
    raw = golden_io.GoldenCase(path).raw
    assert np.array_equal(ctx.twin_match(raw.nv, raw.face_off, np.ascontiguousarray(raw.edges[:, 0])), raw.edges)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(GENERATED.keys()))
def test_cuda_twin_vs_oracle(ctx, name):
    pm = GENERATED[name]()
    want = ol.o_twin_match(pm.nv, pm.face_off, pm.face_idx)
    got = ctx.twin_match(pm.nv, pm.face_off, pm.face_idx)
    assert np.array_equal(got, want)
    # in place on 12-byte records (the layout a reader with Builder::automerge = false holds)
    rec = np.full((pm.face_idx.shape[0], 3), 0xdeadbeef, np.uint32)
    rec[:, 0] = pm.face_idx
    ctx.twin_match(pm.nv, pm.face_off, rec, out=rec)
    assert np.array_equal(rec, want)
    # the slots inside a bucket are handed out by atomics: a second run must give the same table
    assert np.array_equal(ctx.twin_match(pm.nv, pm.face_off, pm.face_idx), want)


@pytest.mark.gpu
def test_cuda_twin_feeds_attr_encode(ctx, workdir):
    """the table hb_twin_match builds is the one the reference's reader hands to the rest of the pipeline"""
    if not ol.have_ref():
        pytest.skip("oracle/_ref/libharry_ref.so not built")
    for name in ("poly_lossless", "irr_q14", "obj_lossless"):
        rm = ol.RefMesh(CASES[name][0](workdir))
        m = rm.arrays()
        rm.close()
        assert np.array_equal(ctx.twin_match(m.nv, m.face_off, np.ascontiguousarray(m.edges[:, 0])), m.edges), name


@pytest.mark.gpu
def test_cuda_twin_empty_and_errors(ctx):
    assert ctx.twin_match(0, np.zeros(1, np.uint32), np.zeros(0, np.uint32)).shape == (0, 3)
    face_off = np.array([0, 3, 3, 6], np.uint32)             # an empty face in the middle
    org = np.array([0, 1, 2, 2, 1, 3], np.uint32)
    assert np.array_equal(ctx.twin_match(4, face_off, org), ol.o_twin_match(4, face_off, org))
    with pytest.raises(capi.HarryError):
        ctx.twin_match(3, face_off, org)                      # vertex 3 >= nv
    n = 65536                                                 # one corner more than fepair::e can address
    with pytest.raises(capi.HarryError):
        ctx.twin_match(n, np.array([0, n], np.uint32), np.arange(n, dtype=np.uint32))
    # the context stays usable after a rejected call
    assert np.array_equal(ctx.twin_match(4, face_off, org), ol.o_twin_match(4, face_off, org))


@pytest.mark.gpu
def test_cuda_twin_1m_vs_oracle(ctx):
    pm = meshgen.uv_sphere(708, 1412)                         # bench.py's CPU sample shape, 6M half-edges
    want = ol.o_twin_match(pm.nv, pm.face_off, pm.face_idx)
    got = ctx.twin_match(pm.nv, pm.face_off, pm.face_idx)
    assert np.array_equal(got, want)
    if ol.have_ref_twin():
        assert np.array_equal(got, ol.ref_twin_match(pm.face_off, pm.face_idx)[0])
    assert check_table_properties(pm.face_off, got) == 0      # closed surface: no border


@pytest.mark.gpu
def test_cuda_twin_full_size_properties(ctx):
    """BASELINE configs[1] connectivity (9 999 394 vertices, 59 996 352 half-edges): properties only"""
    pm = meshgen.uv_sphere(2237, 4472)
    got = ctx.twin_match(pm.nv, pm.face_off, pm.face_idx)
    assert np.array_equal(got[:, 0], pm.face_idx)
    assert check_table_properties(pm.face_off, got) == 0
