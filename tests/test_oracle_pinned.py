"""CPU: pins oracle/harry_oracle.c against the UNMODIFIED reference run live through
oracle/_ref/libharry_ref.so (built from /root/reference).  Covers BASELINE configs 1, 3, 4 at their
small sizes plus quantized / multi-region variants.  Skipped where the reference build is absent."""
import pytest

import checks
import oracle_lib as ol
from cases import CASES, CONFIG1, get_case

pytestmark = pytest.mark.skipif(not ol.have_ref(), reason="oracle/_ref/libharry_ref.so not built (needs /root/reference)")
ALL = list(CASES.keys()) + [CONFIG1[0]]


@pytest.fixture(scope="module")
def impl():
    return checks.OracleImpl()


@pytest.mark.parametrize("name", ALL)
def test_quant(impl, workdir, name):
    checks.check_quant(impl, get_case(workdir, name))


@pytest.mark.parametrize("name", ALL)
def test_encode(impl, workdir, name):
    checks.check_encode(impl, get_case(workdir, name))


@pytest.mark.parametrize("name", ALL)
def test_decode(impl, workdir, name):
    checks.check_decode(impl, get_case(workdir, name))


def test_reference_cli_matches_harness(workdir):
    """the harness' .hry equals what the reference CLI writes for the same input and flags"""
    import os
    import subprocess
    if not os.path.exists(ol.REF_CLI):
        pytest.skip("reference CLI not built")
    case = get_case(workdir, "sphere_q14")
    out = os.path.join(workdir, "cli.hry")
    subprocess.run([ol.REF_CLI, case.src_path, out, "-l1", "-q14"], check=True, capture_output=True)
    assert open(out, "rb").read() == open(case.hry_path, "rb").read()
