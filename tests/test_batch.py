"""GPU parity of the batched path (BASELINE configs[4]: many independent meshes, one launch per stage over all of
them): EVERY mesh of every batch is compared bit for bit with the unmodified reference (oracle/_ref) -- symbol
streams, histograms, bounds rows, quantized / decoded / dequantized rows -- device resident, through the host-buffer
batch entry points (several pipelined groups), and with two contexts working concurrently.  pytest -m gpu."""
import os
import threading

import numpy as np
import pytest

import oracle_lib as ol
from cases import Case, _obj, _ply
from harry_b200 import capi, meshgen

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libharry_ref.so not built")]


def _sphere_family(n):
    fam = []
    for k in range(n):
        nr, ns = 6 + (7 * k) % 31, 7 + (11 * k) % 53
        fam.append((f"bs_{k}", (lambda d, k=k, nr=nr, ns=ns: _ply(d, f"bs_{k}.ply", meshgen.uv_sphere(nr, ns, noise_seed=100 + k))), [(1, -1, 14)]))
    return fam


FAMILIES = {
    # 64 noisy spheres of 64 different sizes (from 44 to 1900 vertices), positions at 14 bits: the shape of configs[4]
    "spheres64_q14": _sphere_family(64),
    "spheres_lossless": [(f"bl_{k}", (lambda d, k=k: _ply(d, f"bl_{k}.ply", meshgen.uv_sphere(9 + 5 * k, 14 + 9 * k, noise_seed=7 + k))), []) for k in range(6)],
    "spheres_q8": [(f"b8_{k}", (lambda d, k=k: _ply(d, f"b8_{k}.ply", meshgen.uv_sphere(12 + 3 * k, 17 + 8 * k, noise_seed=40 + k))), [(1, -1, 8)]) for k in range(5)],
    "spheres_q20": [(f"b20_{k}", (lambda d, k=k: _ply(d, f"b20_{k}.ply", meshgen.uv_sphere(10 + 4 * k, 13 + 6 * k, noise_seed=50 + k))), [(1, -1, 20)]) for k in range(5)],
    # polygon grids with the non-manifold fin: per-vertex and per-face float lists, quantized and not
    "poly_q10": [(f"bp_{k}", (lambda d, k=k: _ply(d, f"bp_{k}.ply", meshgen.poly_grid(8 + 5 * k, seed=3 + k))), [(1, -1, 10), (0, -1, 9)]) for k in range(5)],
    "poly_lossless": [(f"bpl_{k}", (lambda d, k=k: _ply(d, f"bpl_{k}.ply", meshgen.poly_grid(7 + 6 * k, seed=13 + k))), []) for k in range(4)],
    # wide fans (closed and open) and irregular triangulations
    "cones_q12": [(f"bc_{k}", (lambda d, k=k: _ply(d, f"bc_{k}.ply", meshgen.cones(6 + 9 * k, 70 + 30 * k, seed=2 + k, open_every=k % 3))), [(1, -1, 12)]) for k in range(4)],
    "irr_q14": [(f"bi_{k}", (lambda d, k=k: _ply(d, f"bi_{k}.ply", meshgen.tri_irregular(16 + 12 * k, 5 + k))), [(1, -1, 14)]) for k in range(4)],
    # OBJ: corner lists with HIST / LHIST emissions; multi-region variant (several vertex and face regions)
    "obj_q14_q10": [(f"bo_{k}", (lambda d, k=k: _objk(d, f"bo_{k}.obj", 6 + 3 * k, 9 + 4 * k, False)), [(0, -1, 14), (2, -1, 10)]) for k in range(4)],
    # (sizes whose region tables coincide: a batch needs one schema)
    "obj_multi_all": [(f"bom_{k}", (lambda d, k=k, sz=sz: _objk(d, f"bom_{k}.obj", sz[0], sz[1], True)), [(0, -1, 12), (1, -1, 9), (2, -1, 10), (3, -1, 11)])
                      for k, sz in enumerate([(14, 20), (18, 26), (14, 31), (20, 30)])],
}


def _objk(d, name, nr, ns, multi):
    p = os.path.join(d, name)
    if not os.path.exists(p):
        meshgen.write_obj_latlong(p, nr, ns, multi_region=multi)
    return p


_cache = {}


def family(workdir, name):
    if name not in _cache:
        _cache[name] = [Case(workdir, n, gen, loq) for (n, gen, loq) in FAMILIES[name]]
    return _cache[name]


def encoder_inputs(cases):
    """unquantized rows + the reference's traversal order and final twin table"""
    out = []
    for c in cases:
        raw = c.raw.copy()
        raw.order, raw.order_f, raw.edges = c.enc.order, c.enc.order_f, c.enc.edges
        out.append(raw)
    return out


def quant_requests(case):
    return [(l, case.enc.lists[l].quants, la.groups) for l, la in enumerate(case.raw.lists) if la.ncomp and case.enc.lists[l].quants != la.quants]


def check_streams(cases, got):
    assert len(got) == len(cases)
    for i, (c, g) in enumerate(zip(cases, got)):
        ok, why = g.equal(c.enc_streams)
        assert ok, f"mesh {i} ({c.name}): {why}"


def check_bounds(cases, reqs, bounds):
    for (l, _, _), b in zip(reqs, bounds):
        s = cases[0].raw.lists[l].stride
        for i, c in enumerate(cases):
            assert np.array_equal(b[i, 0, :s], c.raw_bounds[l][0]), f"min row, mesh {i} list {l}"
            assert np.array_equal(b[i, 1, :s], c.raw_bounds[l][1]), f"max row, mesh {i} list {l}"
            assert np.array_equal(b[i, 2, :s], c.raw_scale[l]), f"scale row, mesh {i} list {l}"


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("name", list(FAMILIES))
def test_batch_device_resident_encode(ctx, workdir, name):
    cases = family(workdir, name)
    dm = capi.DeviceMesh(ctx, encoder_inputs(cases))
    dm.snapshot()
    reqs = quant_requests(cases[0])
    launches0 = ctx.launches()
    for rep in range(2):
        for l, nq, groups in reqs:
            dm.quantize(l, nq, groups)
        dm.encode()
        check_streams(cases, dm.fetch_streams_batch())
        for l, _, _ in reqs:
            mn, mx, sc = (np.atleast_2d(a) for a in dm.fetch_bounds(l))
            for i, c in enumerate(cases):
                assert np.array_equal(mn[i], c.raw_bounds[l][0]) and np.array_equal(mx[i], c.raw_bounds[l][1]) and np.array_equal(sc[i], c.raw_scale[l]), f"bounds, mesh {i}"
        for i, c in enumerate(cases):
            for l, la in enumerate(c.enc.lists):
                if la.ncomp:
                    assert np.array_equal(dm.fetch_rows(l, i), la.rows), f"quantized rows, mesh {i} list {l}"
        dm.restore()
    # one launch per stage over all meshes: the launch count does not grow with the batch
    assert (ctx.launches() - launches0) / 2 < 150
    dm.close()


@pytest.mark.parametrize("name", list(FAMILIES))
def test_batch_device_resident_decode(ctx, workdir, name):
    cases = family(workdir, name)
    dm = capi.DeviceMesh(ctx, [c.decode_input() for c in cases])
    for l, la in enumerate(cases[0].dec.lists):
        if la.ncomp:
            sc = np.stack([c.deq_scale[l] for c in cases]) if cases[0].deq is not None else None
            dm.set_bounds(l, np.stack([c.dec_bounds[l][0] for c in cases]), np.stack([c.dec_bounds[l][1] for c in cases]), sc)
    dm.snapshot()
    for rep in range(2):
        dm.decode()
        for i, c in enumerate(cases):
            for l, la in enumerate(c.dec.lists):
                if la.ncomp:
                    assert np.array_equal(dm.fetch_rows(l, i), la.rows), f"rep {rep}: decoded rows, mesh {i} ({c.name}) list {l}"
        if cases[0].deq is not None:
            for l, la in enumerate(cases[0].dec.lists):
                if any(la.quants):
                    dm.dequantize(l)
                    for i, c in enumerate(cases):
                        assert np.array_equal(dm.fetch_rows(l, i), c.deq.lists[l].rows), f"dequantized rows, mesh {i} list {l}"
        dm.restore()
    dm.close()


@pytest.mark.parametrize("ntb,dense", [("128", "1"), ("128", "0"), ("256", "0")])
@pytest.mark.parametrize("name", ["spheres64_q14", "obj_multi_all"])
def test_batch_decode_scan_variants(ctx, workdir, name, ntb, dense, monkeypatch):
    """The scan decoder picks its CTA size and register budget from the number of chains (128 threads with 8 CTAs per SM
    from four chains per SM on): every variant reconstructs the same rows, whatever the batch size that selects it"""
    monkeypatch.setenv("HARRY_B200_SCAN_NTB", ntb)
    monkeypatch.setenv("HARRY_B200_SCAN_DENSE", dense)
    test_batch_device_resident_decode(ctx, workdir, name)


_pinned_keep = []


def page_locked(mesh, salt):
    """the mesh with every array in page-locked host memory (torch is only the allocator), at varying offsets from a
    16-byte boundary: what the gathered upload of a batch reads directly from the device"""
    import torch

    def pin(a, k):
        a = np.ascontiguousarray(a)
        off = ((salt + k) % 4) * a.dtype.itemsize if a.dtype.itemsize < 16 else 0   # 0, 1, 2 or 3 elements off the boundary
        buf = torch.empty(a.nbytes + 64, dtype=torch.uint8, pin_memory=True)
        _pinned_keep.append(buf)
        out = buf.numpy()[off: off + a.nbytes].view(a.dtype).reshape(a.shape)
        out[...] = a
        return out

    m = mesh.copy()
    for k, name in enumerate(("edges", "face_off", "order", "order_f", "vtx_regs", "face_regs", "bind_face", "bind_vtx", "bind_corner")):
        a = getattr(m, name)
        if a is not None and a.size:
            setattr(m, name, pin(a, k))
    for k, la in enumerate(m.lists):
        if la.rows.size:
            la.rows = pin(la.rows, 9 + k)
    if m.emit_types is not None:
        m.emit_types = [pin(t, 13 + k) if t is not None and t.size else t for k, t in enumerate(m.emit_types)]
    return m


def host_roundtrip(ctx, cases, pinned=False):
    reqs = quant_requests(cases[0])
    enc_in = encoder_inputs(cases)
    if pinned:
        enc_in = [page_locked(m, i) for i, m in enumerate(enc_in)]
    streams, bounds = ctx.encode_batch(enc_in, reqs)
    check_streams(cases, streams)
    check_bounds(cases, reqs, bounds)
    ins = [c.decode_input() for c in cases]
    if pinned:
        ins = [page_locked(m, 3 + i) for i, m in enumerate(ins)]
    deq = []
    if cases[0].deq is not None:
        for l, la in enumerate(cases[0].dec.lists):
            if any(la.quants):
                deq.append((l, np.stack([np.stack([c.dec_bounds[l][0], c.dec_bounds[l][1], c.deq_scale[l]]) for c in cases])))
    ctx.decode_batch(ins, deq)
    for i, (c, m) in enumerate(zip(cases, ins)):
        want = c.deq if c.deq is not None else c.dec
        for l, la in enumerate(want.lists):
            if la.ncomp:
                assert np.array_equal(m.lists[l].rows, la.rows), f"decoded rows, mesh {i} ({c.name}) list {l}"


@pytest.mark.parametrize("name", list(FAMILIES))
@pytest.mark.parametrize("group", [0, 9000])
def test_batch_host_buffers(ctx, workdir, name, group, monkeypatch):
    """hb_encode_batch / hb_decode_batch: one group, and many small pipelined groups (a few meshes each)"""
    if group:
        monkeypatch.setenv("HARRY_B200_GROUP_HALF_EDGES", str(group))
    host_roundtrip(ctx, family(workdir, name))


@pytest.mark.parametrize("name", ["spheres64_q14", "poly_q10", "obj_multi_all", "spheres_lossless"])
@pytest.mark.parametrize("group", [0, 20000])
@pytest.mark.parametrize("mode", ["2", "1", "0"])
def test_batch_host_buffers_page_locked(ctx, workdir, name, group, mode, monkeypatch):
    """page-locked host arrays of a batch are not copied one by one: they are queued and issued per upload stage as ONE
    batched copy (mode 2, cudaMemcpyBatchAsync) or fetched by one kernel that reads the host memory directly (mode 1: any
    alignment of source and destination); mode 0 = one copy per array.  Same results every way."""
    monkeypatch.setenv("HARRY_B200_GATHER_UPLOADS", mode)
    if group:
        monkeypatch.setenv("HARRY_B200_GROUP_HALF_EDGES", str(group))
    launches0, up0 = ctx.launches(), ctx.h2d_bytes()
    host_roundtrip(ctx, family(workdir, name), pinned=True)
    assert ctx.h2d_bytes() > up0
    _pinned_keep.clear()


def test_batch_two_contexts_concurrently(workdir, monkeypatch):
    """two host threads, each with its own context (stream set, memory pool use, page-locked cache are shared or
    per context): every mesh of both batches is still the reference's, over several rounds"""
    monkeypatch.setenv("HARRY_B200_GROUP_HALF_EDGES", "30000")
    fams = [family(workdir, "spheres64_q14"), family(workdir, "obj_multi_all") + [], family(workdir, "poly_q10")]
    errors = []

    def worker(k):
        try:
            c = capi.Context(0)
            for rnd in range(3):
                host_roundtrip(c, fams[0][k::2] if rnd != 1 else fams[1 + k])
            c.close()
        except BaseException as e:  # noqa: BLE001 -- reported by the main thread
            errors.append((k, repr(e)))

    ts = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors


def test_batch_rejects_mixed_schemas(ctx, workdir):
    a = family(workdir, "spheres_q8")[0]
    b = family(workdir, "poly_q10")[0]
    with pytest.raises(capi.HarryError):
        capi.DeviceMesh(ctx, encoder_inputs([a, b]))
