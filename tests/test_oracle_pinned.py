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


@pytest.mark.parametrize("name", ["sphere_q14", "sphere_lossless", "rgb_xyz_q12", "rgb_q5_xyz_q14"])
def test_face_order_is_irrelevant_without_face_attributes(workdir, name):
    """The premise of the encoder's upload shortcut, on the CPU restatement: with no FACE / CORNER
    components, one face region and no face row bound twice, the streams of the reference's traversal order of the faces
    are the streams of the index order with gate corner 0 -- which is what the library codes when it leaves order_f on
    the host."""
    import numpy as np
    case = get_case(workdir, name)
    mesh = case.enc
    assert len(mesh.off_reg_face) == 2 and all(la.ncomp == 0 or la.target == 1 for la in mesh.lists)
    assert len(np.unique(mesh.bind_face)) == mesh.bind_face.size          # no face row bound twice
    assert not np.array_equal(mesh.order_f[:, 0], np.arange(mesh.nf))     # the traversal is not the index order
    plain = mesh.copy()
    plain.order_f = np.stack([np.arange(mesh.nf, dtype=np.uint32), np.zeros(mesh.nf, dtype=np.uint32)], axis=1)
    ok, why = ol.o_attr_encode(plain).equal(case.enc_streams)
    assert ok, why
    none = mesh.copy()
    none.order_f = None
    ok, why = ol.o_attr_encode(none).equal(case.enc_streams)
    assert ok, why
