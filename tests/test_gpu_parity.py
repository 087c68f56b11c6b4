"""GPU parity: the CUDA path (through the C ABI) against the real reference's captured outputs and
the CPU oracle, bit for bit.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import oracle_lib as ol
from cases import CASES, CONFIG1, get_case
from harry_b200 import capi

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libharry_ref.so not built")


@pytest.fixture(scope="module")
def ctx():
    c = capi.Context(0)
    yield c
    c.close()


ALL = list(CASES.keys()) + [CONFIG1[0]]


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_bounds_and_requant(ctx, workdir, name):
    case = get_case(workdir, name)
    for l, la in enumerate(case.raw.lists):
        if la.ncomp == 0:
            continue
        mn, mx = ctx.bounds(la)
        assert np.array_equal(mn, case.raw_bounds[l][0]), f"min row list {l}"
        assert np.array_equal(mx, case.raw_bounds[l][1]), f"max row list {l}"
        nq = case.enc.lists[l].quants
        if nq != la.quants:
            lb = la.copy()
            ctx.requant(lb, nq, mn, case.raw_scale[l])
            assert lb.quants == nq
            assert np.array_equal(lb.rows, case.enc.lists[l].rows), f"quantized rows list {l}"


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_encode_streams(ctx, workdir, name):
    case = get_case(workdir, name)
    got = ctx.attr_encode(case.enc)
    ok, why = got.equal(case.enc_streams)
    assert ok, why
    # and the CPU restatement agrees as well
    ok, why = got.equal(ol.o_attr_encode(case.enc))
    assert ok, why
    assert ctx.launches() > 0


@needs_ref
@pytest.mark.parametrize("name", ALL)
def test_decode_rows(ctx, workdir, name):
    case = get_case(workdir, name)
    m = case.decode_input()
    ctx.attr_decode(m)
    for l, la in enumerate(case.dec.lists):
        assert np.array_equal(m.lists[l].rows, la.rows), f"decoded rows list {l}"
    if case.deq is not None:
        for l, la in enumerate(case.dec.lists):
            if not any(la.quants):
                continue
            lb = m.lists[l].copy()
            ctx.requant(lb, [0] * la.ncomp, case.dec_bounds[l][0], case.deq_scale[l])
            assert np.array_equal(lb.rows, case.deq.lists[l].rows), f"dequantized rows list {l}"


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
