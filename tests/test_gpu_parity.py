"""GPU parity: the CUDA path, called through the C ABI, against (a) the committed golden vectors,
(b) the unmodified reference run live through oracle/_ref, and (c) the CPU oracle -- bit for bit.
Run on the B200 box: pytest -m gpu."""
import glob
import os

import numpy as np
import pytest

import checks
import golden_io
import oracle_lib as ol
from cases import CASES, CONFIG1, get_case
from harry_b200 import capi

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libharry_ref.so not built")
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
GIDS = [os.path.basename(p)[:-4] for p in GOLDEN]
ALL = list(CASES.keys()) + [CONFIG1[0]]


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def impl(ctx):
    return checks.CudaImpl(ctx)


# ---- committed golden vectors -----------------------------------------------------------------
@pytest.mark.parametrize("path", GOLDEN, ids=GIDS)
def test_golden_quant(impl, path):
    checks.check_quant(impl, golden_io.GoldenCase(path))


@pytest.mark.parametrize("path", GOLDEN, ids=GIDS)
def test_golden_encode(impl, path):
    case = golden_io.GoldenCase(path)
    got = checks.check_encode(impl, case)
    ok, why = got.equal(ol.o_attr_encode(case.enc))
    assert ok, why


@pytest.mark.parametrize("path", GOLDEN, ids=GIDS)
def test_golden_decode(impl, path):
    checks.check_decode(impl, golden_io.GoldenCase(path))


# ---- live reference -----------------------------------------------------------------------------
@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_bounds_and_requant(impl, workdir, name):
    checks.check_quant(impl, get_case(workdir, name))


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_encode_streams(impl, ctx, workdir, name):
    checks.check_encode(impl, get_case(workdir, name))
    assert ctx.launches() > 0


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_decode_rows(impl, workdir, name):
    checks.check_decode(impl, get_case(workdir, name))


@needs_ref
@pytest.mark.parametrize("name", ["sphere_q14", "obj_multi_all", "poly_q10"])
def test_device_resident_pipeline(ctx, workdir, name):
    """upload once -> quantize -> encode on the device; same streams as the host-buffer path."""
    case = get_case(workdir, name)
    raw = case.raw.copy()
    raw.order, raw.order_f, raw.edges = case.enc.order, case.enc.order_f, case.enc.edges
    dm = capi.DeviceMesh(ctx, raw)
    dm.snapshot()
    for rep in range(2):
        for l, la in enumerate(raw.lists):
            nq = case.enc.lists[l].quants
            if la.ncomp and nq != la.quants:
                dm.quantize(l, nq, la.groups)
        dm.encode()
        got = dm.fetch_streams()
        ok, why = got.equal(case.enc_streams)
        assert ok, f"rep {rep}: {why}"
        for l, la in enumerate(raw.lists):
            if la.ncomp:
                assert np.array_equal(dm.fetch_rows(l), case.enc.lists[l].rows)
        dm.restore()
    dm.close()


@needs_ref
@pytest.mark.parametrize("name", ["sphere_q14", "sphere_lossless", "obj_q14_q10"])
def test_device_resident_decode(ctx, workdir, name):
    """decode twice from a device snapshot (the path bench.py times)"""
    case = get_case(workdir, name)
    m = case.decode_input()
    dm = capi.DeviceMesh(ctx, m)
    for l, (mn, mx) in enumerate(case.dec_bounds):
        if m.lists[l].ncomp:
            dm.set_bounds(l, mn, mx, case.deq_scale[l] if case.deq is not None else None)
    dm.snapshot()
    for rep in range(2):
        dm.decode()
        for l, la in enumerate(case.dec.lists):
            if la.ncomp:
                assert np.array_equal(dm.fetch_rows(l), la.rows), f"rep {rep} list {l}"
        if case.deq is not None:
            for l, la in enumerate(case.dec.lists):
                if any(la.quants):
                    dm.dequantize(l)
                    assert np.array_equal(dm.fetch_rows(l), case.deq.lists[l].rows)
        dm.restore()
    dm.close()


# ---- size-independent properties at larger sizes ------------------------------------------------
def test_roundtrip_property_large(ctx):
    """encode -> decode on the same (encoder-numbered) mesh is the identity for the quantized values,
    at a size where the oracle still finishes in seconds; the oracle streams agree too."""
    from harry_b200 import flatten, meshgen
    pm = meshgen.uv_sphere(300, 601, noise_seed=9)           # 179 501 vertices
    mesh = flatten.mesh_arrays(pm)
    vl = mesh.lists[1]
    mn, mx = ctx.bounds(vl)
    sc = ol.o_scale(vl, mn, mx)
    ctx.requant(vl, [14, 14, 14], mn, sc)
    got = ctx.attr_encode(mesh)
    ok, why = got.equal(ol.o_attr_encode(mesh))
    assert ok, why
    assert int(got.lists[1].hist.sum()) == 6 * mesh.nv        # checksum of checksums: one count per symbol
    dec = mesh.copy()
    dec.lists = capi.residual_rows_encoder_side(mesh, got)
    ctx.attr_decode(dec)
    cols = [0, 1, 4, 5, 8, 9]
    assert np.array_equal(dec.lists[1].rows[:, cols], mesh.lists[1].rows[:, cols])


# ---- BASELINE configs[1] at its full size -----------------------------------------------------------
@needs_ref
def test_config2_full_size(ctx, workdir):
    """The 10M-vertex mesh bench.py measures: encode streams and decoded rows equal the reference's own
    AttrCoder / AttrDecoder output (about a minute, most of it the reference preparing the inputs)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    w = bench.Workload(*bench.FULL, workdir, keep_expected=True)
    E = capi.DeviceMesh(ctx, w.raw)
    E.quantize(1, w.new_quant[1], w.raw.lists[1].groups)
    E.encode()
    ok, why = E.fetch_streams().equal(w.enc_expected)
    E.close()
    assert ok, why
    D = capi.DeviceMesh(ctx, w.dec)
    for l, (mn, mx, sc) in enumerate(w.dec_bounds):
        if w.dec.lists[l].ncomp:
            D.set_bounds(l, mn, mx, sc)
    D.decode()
    for l, exp in enumerate(w.dec_expected):
        if w.dec.lists[l].ncomp:
            assert np.array_equal(D.fetch_rows(l), exp), f"list {l}"
    D.close()


# ---- BASELINE configs[2] and [3] at scale (corner lists with thousands of LHIST hits per region, deep polygon meshes) ----
def _full_case_check(impl, workdir, name, gen, loq):
    from cases import Case
    case = Case(workdir, name, gen, loq)
    checks.check_quant(impl, case)
    checks.check_encode(impl, case)
    checks.check_decode(impl, case)
    return case


@needs_ref
def test_config3_at_scale(impl, workdir):
    """OBJ lat-long sphere, 240 600 vertices / 480 000 triangles / 1.44 M corners, vt + vn corner lists, -l0 -q14 -l2 -q10"""
    from harry_b200 import meshgen

    def gen(d):
        p = os.path.join(d, "cfg3_scale.obj")
        meshgen.write_obj_latlong(p, 400, 600)
        return p

    case = _full_case_check(impl, workdir, "cfg3_scale", gen, [(0, -1, 14), (2, -1, 10)])
    assert case.enc.nv == 240600 and int((case.enc_streams.lists[1].type == capi.LHIST).sum()) > 1000000


@needs_ref
def test_config3_multi_region_at_scale(impl, workdir):
    from harry_b200 import meshgen

    def gen(d):
        p = os.path.join(d, "cfg3m_scale.obj")
        meshgen.write_obj_latlong(p, 150, 220, multi_region=True)
        return p

    _full_case_check(impl, workdir, "cfg3m_scale", gen, [(0, -1, 12), (1, -1, 9), (2, -1, 10), (3, -1, 11)])


@needs_ref
@pytest.mark.parametrize("loq", [[], [(1, -1, 12), (0, -1, 9)]], ids=["lossless", "q12_q9"])
def test_config4_at_scale(impl, workdir, loq):
    """polygon grid n = 400 (161 607 vertices; tri / quad / 5- / 6-gons + the non-manifold fin), per-vertex and per-face floats"""
    from cases import _ply
    from harry_b200 import meshgen
    _full_case_check(impl, workdir, "cfg4_scale" + ("q" if loq else ""), lambda d: _ply(d, "cfg4_scale.ply", meshgen.poly_grid(400)), loq)


@needs_ref
@pytest.mark.parametrize("loq", [[(1, -1, 12)], []], ids=["q12", "lossless"])
def test_many_wide_fans(impl, workdir, loq):
    """4 300 fans of 70 faces: more wide vertices than the wide-fan path takes at once (4096 slots) and than the encode
    kernels defer to their cooperative path (wide_cap): the rest is walked / summed by single threads -- same streams"""
    from cases import _ply
    from harry_b200 import meshgen
    _full_case_check(impl, workdir, "many_cones" + ("q" if loq else ""), lambda d: _ply(d, "many_cones.ply", meshgen.cones(4300, 70, seed=3)), loq)


def test_zero_component_list_with_shared_rows(ctx):
    """A list without components still emits DATA / HIST type symbols and history offsets (io.h:90-94,
    attrcode.h:43-52).  No reader of the reference builds shared rows for such a list, so the two-pass path is
    checked against the oracle on hand-made bindings: pairs of faces share a row, in a batch of two meshes."""
    from harry_b200 import flatten, meshgen
    meshes = []
    for seed, (nr, ns) in ((3, (9, 14)), (5, (12, 11))):
        mesh = flatten.mesh_arrays(meshgen.uv_sphere(nr, ns, noise_seed=seed))
        fl = mesh.lists[0]
        assert fl.ncomp == 0 and fl.target == capi.T_FACE
        perm = np.random.default_rng(seed).permutation(mesh.nf)
        mesh.bind_face = (perm // 2).astype(np.uint32)                      # two faces per row, scattered over the traversal
        mesh.lists[0] = capi.empty_list((mesh.nf + 1) // 2, capi.T_FACE)
        meshes.append(mesh)
        got = ctx.attr_encode(mesh)
        want = ol.o_attr_encode(mesh)
        ok, why = got.equal(want)
        assert ok, why
        assert int((got.lists[0].type == capi.HIST).sum()) == mesh.nf // 2
    dm = capi.DeviceMesh(ctx, meshes)
    dm.encode()
    for mesh, got in zip(meshes, dm.fetch_streams_batch()):
        ok, why = got.equal(ol.o_attr_encode(mesh))
        assert ok, why
    dm.close()
    # host-buffer batch: the encoder uploads the face order only after the device found a row that is bound twice
    up0 = ctx.h2d_bytes()
    got_b, _ = ctx.encode_batch(meshes)
    assert ctx.h2d_bytes() - up0 >= sum(m.order_f.nbytes for m in meshes)
    for mesh, got in zip(meshes, got_b):
        ok, why = got.equal(ol.o_attr_encode(mesh))
        assert ok, why


@needs_ref
@pytest.mark.parametrize("name", ["sphere_q14", "poly_q10", "obj_multi_all"])
def test_face_order_not_uploaded_when_unneeded(workdir, name, monkeypatch):
    """hb_attr_encode / hb_encode_batch leave the face order (8 bytes per face) on the host when no FACE / CORNER list
    carries components, the mesh has one face region and the device finds no face row bound twice: same streams as with
    the order (HARRY_B200_KEEP_ORDER_F=1), fewer bytes on the link.  Meshes with face attributes keep uploading it."""
    case = get_case(workdir, name)
    c = capi.Context(0)
    mesh = case.enc
    droppable = mesh.order_f is not None and len(mesh.off_reg_face) == 2 and mesh.order_f.shape[0] == mesh.nf and \
        all(la.ncomp == 0 or la.target == capi.T_VTX for la in mesh.lists)
    up = []
    for keep in ("1", "0"):
        monkeypatch.setenv("HARRY_B200_KEEP_ORDER_F", keep)
        u0 = c.h2d_bytes()
        got = c.attr_encode(mesh)
        ok, why = got.equal(case.enc_streams)
        assert ok, f"keep={keep}: {why}"
        got_b, _ = c.encode_batch([mesh, mesh])
        up.append(c.h2d_bytes() - u0)
        for g in got_b:
            ok, why = g.equal(case.enc_streams)
            assert ok, f"batch, keep={keep}: {why}"
    # one mesh through hb_attr_encode, two through hb_encode_batch; the check itself uploads a few words (row-bitmap bases)
    saved, want = up[0] - up[1], (3 * mesh.order_f.nbytes if droppable else 0)
    assert want - 64 <= saved <= want, (saved, want)
    c.close()


@needs_ref
@pytest.mark.parametrize("name", ["sphere_q14", "sphere_lossless", "obj_multi_all", "rgb_q5_xyz_q14", "poly_q10"])
def test_row_cache_pipeline(workdir, name):
    """hb_ctx_set_row_cache: the adapter's call order on the SAME host buffers (set_bounds -> requant -> encode, decode ->
    requant(clear)), twice over refilled buffers -- same results as without the cache, no stale device rows"""
    case = get_case(workdir, name)
    c = capi.Context(0)
    c.set_row_cache(True)
    raw = case.raw.copy()
    raw.order, raw.order_f, raw.edges = case.enc.order, case.enc.order_f, case.enc.edges
    m = case.decode_input()
    fresh_raw = [la.rows.copy() for la in raw.lists]
    fresh_dec = [la.rows.copy() for la in m.lists]
    dec_quants = [list(la.quants) for la in m.lists]
    for rep in range(2):
        for l, la in enumerate(raw.lists):
            la.rows[...] = fresh_raw[l]
            la.quants = [0] * la.ncomp
            if not la.ncomp:
                continue
            mn, mx = c.bounds(la)
            assert np.array_equal(mn, case.raw_bounds[l][0]) and np.array_equal(mx, case.raw_bounds[l][1])
            nq = case.enc.lists[l].quants
            if nq != la.quants:
                c.requant(la, nq, mn, case.raw_scale[l])
                assert np.array_equal(la.rows, case.enc.lists[l].rows), f"rep {rep}: quantized rows, list {l}"
        ok, why = c.attr_encode(raw).equal(case.enc_streams)
        assert ok, f"rep {rep}: {why}"
        for l, la in enumerate(m.lists):
            la.rows[...] = fresh_dec[l]
            la.quants = list(dec_quants[l])
        c.attr_decode(m)
        for l, la in enumerate(case.dec.lists):
            assert np.array_equal(m.lists[l].rows, la.rows), f"rep {rep}: decoded rows, list {l}"
        if case.deq is not None:
            for l, la in enumerate(case.dec.lists):
                if any(la.quants):
                    c.requant(m.lists[l], [0] * la.ncomp, case.dec_bounds[l][0], case.deq_scale[l])
                    assert np.array_equal(m.lists[l].rows, case.deq.lists[l].rows), f"rep {rep}: dequantized rows, list {l}"
    c.set_row_cache(False)
    c.close()


# ---- malformed input: an error code, never an out-of-bounds access; the context stays usable ----------------------------
def _corruptions():
    def bad_org(m):
        m.edges[5, 0] = m.nv + 7

    def bad_twin_face(m):
        m.edges[9, 1] = m.nf + 3

    def bad_twin_edge(m):
        m.edges[11, 2] = 200

    def bad_order(m):
        m.order[4, 0] = m.nf + 1

    def bad_order_edge(m):
        m.order[7, 1] = 9

    def vertex_twice(m):
        m.order[10] = m.order[3]        # a vertex listed twice in the traversal order (candidate bound 2 ne still holds or is flagged)

    def bad_binding(m):
        m.bind_vtx[2] = m.lists[1].nrows + 5

    def bad_face_offsets(m):
        m.face_off[3] = m.face_off[5] + 2   # not monotone: faces 2 / 3 have an impossible extent

    return [bad_org, bad_twin_face, bad_twin_edge, bad_order, bad_order_edge, vertex_twice, bad_binding, bad_face_offsets]


@pytest.mark.parametrize("corrupt", _corruptions(), ids=lambda f: f.__name__)
def test_invalid_input_is_reported(corrupt):
    from harry_b200 import flatten, meshgen
    c = capi.Context(0)
    good = flatten.mesh_arrays(meshgen.uv_sphere(14, 19, noise_seed=2))
    want = c.attr_encode(good)
    for batch in (False, True):
        bad = good.copy()
        for name in ("edges", "order", "bind_vtx", "face_off"):
            setattr(bad, name, getattr(bad, name).copy())
        corrupt(bad)
        try:
            if batch:
                c.encode_batch([good, bad, good])
            else:
                c.attr_encode(bad)
            ok = corrupt.__name__ == "vertex_twice"      # legal for the coder as long as the candidate bound holds
            assert ok, "malformed mesh was accepted"
        except capi.HarryError as e:
            assert "(-2)" in str(e) or "(-3)" in str(e), str(e)
        # the same context still produces the right streams afterwards
        ok, why = c.attr_encode(good).equal(want)
        assert ok, f"context unusable after {corrupt.__name__}: {why}"
        bad2 = good.copy()
        bad2.lists = capi.residual_rows_encoder_side(good, want)
        for name in ("edges", "order", "bind_vtx", "face_off"):
            setattr(bad2, name, getattr(bad2, name).copy())
        corrupt(bad2)
        try:
            c.attr_decode(bad2)
        except capi.HarryError:
            pass
    c.close()


def _emptied(mesh):
    """the same schema with no vertex, no face and no row"""
    e = mesh.copy()
    e.nv = e.nf = 0
    e.edges = np.zeros((0, 3), np.uint32)
    e.face_off = np.zeros(1, np.uint32)
    e.order = np.zeros((0, 2), np.uint32)
    e.order_f = np.zeros((0, 2), np.uint32) if mesh.order_f is not None else None
    e.vtx_regs = np.zeros(0, np.uint16)
    e.face_regs = np.zeros(0, np.uint16)
    e.bind_face = np.zeros(0, np.uint32)
    e.bind_vtx = np.zeros(0, np.uint32)
    e.bind_corner = np.zeros(0, np.uint32)
    for la in e.lists:
        la.rows = np.zeros((0, la.rows.shape[1]), np.uint8)
    return e


@needs_ref
@pytest.mark.parametrize("name", ["sphere_q14", "poly_q10"])
def test_empty_mesh_alone_and_inside_a_batch(workdir, name):
    """no vertex, no face, no row: empty streams (like the oracle's), and an empty mesh between two others leaves
    their streams untouched (empty segments in every per-segment table)"""
    case = get_case(workdir, name)
    c = capi.Context(0)
    full = case.enc
    empty = _emptied(full)
    want_empty = ol.o_attr_encode(empty)
    ok, why = c.attr_encode(empty).equal(want_empty)
    assert ok, why
    got, _ = c.encode_batch([full, empty, full, empty])
    for g, want in zip(got, (case.enc_streams, want_empty, case.enc_streams, want_empty)):
        ok, why = g.equal(want)
        assert ok, why
    dm = capi.DeviceMesh(c, [empty, full, empty])
    dm.encode()
    for g, want in zip(dm.fetch_streams_batch(), (want_empty, case.enc_streams, want_empty)):
        ok, why = g.equal(want)
        assert ok, why
    dm.close()
    # decode side: the empty decoder mesh between two real ones
    ins = [case.decode_input(), _emptied(case.decode_input()), case.decode_input()]
    ins[1].emit_types = [np.zeros(0, np.uint8) for _ in ins[1].lists]
    c.decode_batch(ins)
    for m in (ins[0], ins[2]):
        for l, la in enumerate(case.dec.lists):
            if la.ncomp:
                assert np.array_equal(m.lists[l].rows, la.rows)
    c.close()
