"""CPU: the oracle restatement against the committed golden vectors (generated from the unmodified
reference by tests/golden/make_golden.py).  Needs neither the reference nor a GPU."""
import glob
import os

import pytest

import checks
import golden_io

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


@pytest.fixture(scope="module")
def impl():
    return checks.OracleImpl()


def test_fixtures_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_quant(impl, path):
    checks.check_quant(impl, golden_io.GoldenCase(path))


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_encode(impl, path):
    checks.check_encode(impl, golden_io.GoldenCase(path))


@pytest.mark.parametrize("path", GOLDEN, ids=IDS)
def test_decode(impl, path):
    checks.check_decode(impl, golden_io.GoldenCase(path))
