"""CPU: scalar properties of the oracle's residual arithmetic (formats/hry/prediction.h,
transform.h, arith/msb.h as restated in oracle/harry_oracle.c)."""
import numpy as np
import pytest

import oracle_lib as ol
from harry_b200 import capi


@pytest.mark.parametrize("stype,q", [(capi.UCHAR, 3), (capi.UCHAR, 6), (capi.UCHAR, 8), (capi.USHORT, 10)])
def test_decode_inverts_encode_exhaustively(stype, q):
    lib = ol.oracle()
    top = 1 << q
    for pred in range(top):
        for raw in range(0, top, 1 if q <= 8 else 7):
            d = lib.ho_encode_delta(stype, raw, pred, q)
            assert d < top, "residual codes never exceed mask(bits)"
            assert lib.ho_decode_delta(stype, d, pred, q) == raw


def test_decode_inverts_encode_sampled_q14_and_float():
    lib = ol.oracle()
    rng = np.random.default_rng(0)
    for raw, pred in rng.integers(0, 1 << 14, size=(20000, 2)):
        d = lib.ho_encode_delta(capi.USHORT, int(raw), int(pred), 14)
        assert d < (1 << 14)
        assert lib.ho_decode_delta(capi.USHORT, d, int(pred), 14) == raw
    vals = np.concatenate([rng.standard_normal(2000).astype(np.float32), np.array([0.0, -0.0, 1e-38, -1e38, 3.4e38], np.float32)])
    bits = vals.view(np.uint32)
    for a in bits[:300]:
        for b in bits[-40:]:
            d = lib.ho_encode_delta(capi.FLOAT, int(a), int(b), 0)
            assert lib.ho_decode_delta(capi.FLOAT, d, int(b), 0) == int(a)


def test_predict_saturates():
    lib = ol.oracle()
    m = (1 << 14) - 1
    assert lib.ho_predict(capi.USHORT, 16000, 16000, 100, 14) == m      # overflow -> max
    assert lib.ho_predict(capi.USHORT, 10, 100, 16000, 14) == 0         # underflow -> 0
    assert lib.ho_predict(capi.USHORT, 100, 50, 20, 14) == 130
    f = lambda x: int(np.float32(x).view(np.uint32))
    assert lib.ho_predict(capi.FLOAT, f(1.5), f(2.25), f(0.5), 0) == f(np.float32(1.5) + (np.float32(2.25) - np.float32(0.5)))


def test_msb():
    lib = ol.oracle()
    for x, want in [(1, 1), (2, 2), (3, 2), (255, 128), (256, 256), (0x80000001, 0x80000000), (0, 0)]:
        assert lib.ho_msb(x) == want
