"""Parity assertions shared by the oracle (CPU) and CUDA (GPU) test modules.  `impl` is an object
with bounds / requant / attr_encode / attr_decode taking capi containers; `case` is a cases.Case
(live reference output) or a golden_io.GoldenCase (committed reference output)."""
from __future__ import annotations

import numpy as np

import oracle_lib as ol


class OracleImpl:
    """oracle/libharry_oracle.so -- the CPU restatement."""
    name = "oracle"

    def bounds(self, la):
        return ol.o_bounds(la)

    def requant(self, la, nq, mn, sc):
        ol.o_requant(la, nq, mn, sc)

    def attr_encode(self, mesh):
        return ol.o_attr_encode(mesh)

    def attr_decode(self, mesh):
        ol.o_attr_decode(mesh)


class CudaImpl:
    """harry_b200/libharry_b200.so through the C ABI."""
    name = "cuda"

    def __init__(self, ctx):
        self.ctx = ctx

    def bounds(self, la):
        return self.ctx.bounds(la)

    def requant(self, la, nq, mn, sc):
        self.ctx.requant(la, nq, mn, sc)

    def attr_encode(self, mesh):
        return self.ctx.attr_encode(mesh)

    def attr_decode(self, mesh):
        self.ctx.attr_decode(mesh)


def check_quant(impl, case):
    """set_bounds rows and quantized rows equal the reference's, byte for byte."""
    for l, la in enumerate(case.raw.lists):
        if la.ncomp == 0:
            continue
        mn, mx = impl.bounds(la)
        assert np.array_equal(mn, case.raw_bounds[l][0]), f"min row, list {l}"
        assert np.array_equal(mx, case.raw_bounds[l][1]), f"max row, list {l}"
        if impl.name == "oracle":
            assert np.array_equal(ol.o_scale(la, mn, mx), case.raw_scale[l]), f"scale row, list {l}"
        nq = case.enc.lists[l].quants
        if nq != la.quants:
            lb = la.copy()
            impl.requant(lb, nq, mn, case.raw_scale[l])
            assert lb.quants == nq
            assert np.array_equal(lb.rows, case.enc.lists[l].rows), f"quantized rows, list {l}"


def check_encode(impl, case):
    """region / type / history-offset / residual-symbol streams and the per-context histograms equal
    what the reference's AttrCoder fed its writer (and its models' final frequency tables)."""
    got = impl.attr_encode(case.enc)
    ok, why = got.equal(case.enc_streams)
    assert ok, why
    return got


def check_decode(impl, case):
    """rows reconstructed from the residual rows the reference's decoder read equal its output;
    requant(clear) of them equals the reference's dequantized rows."""
    m = case.decode_input()
    impl.attr_decode(m)
    for l, la in enumerate(case.dec.lists):
        assert np.array_equal(m.lists[l].rows, la.rows), f"decoded rows, list {l}"
    if case.deq is not None:
        for l, la in enumerate(case.dec.lists):
            if not any(la.quants):
                continue
            lb = m.lists[l].copy()
            impl.requant(lb, [0] * la.ncomp, case.dec_bounds[l][0], case.deq_scale[l])
            assert np.array_equal(lb.rows, case.deq.lists[l].rows), f"dequantized rows, list {l}"
    # without the drained type symbols the first-reference rule must give the same result whenever
    # no LHIST emission precedes the DATA emission of its row
    m2 = case.decode_input()
    quirk = False
    for l, ls in enumerate(case.dec_streams.lists):
        if (ls.type == 2).any():
            quirk = True
    if not quirk:
        m2.emit_types = None
        impl.attr_decode(m2)
        for l, la in enumerate(case.dec.lists):
            assert np.array_equal(m2.lists[l].rows, la.rows), f"decoded rows (no type stream), list {l}"
