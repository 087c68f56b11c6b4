"""CPU: the C-ABI library loads and exports every symbol include/harry_b200.h declares; the ctypes
struct mirrors have the C sizes; without a GPU the entry points fail loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import pytest

from harry_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "harry_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_python_symbol_lists_agree():
    assert set(declared_functions()) == set(capi.EXPORTED_SYMBOLS)


@pytest.mark.skipif(not os.path.exists(capi.LIB_PATH), reason="libharry_b200.so not built")
def test_library_exports_every_declared_symbol():
    lib = C.CDLL(capi.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} missing from libharry_b200.so"


def test_struct_sizes_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "harry_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(hb_list_desc), sizeof(hb_mesh_desc), sizeof(hb_list_streams), sizeof(hb_streams), sizeof(hb_batch_streams), sizeof(hb_quant_req), sizeof(hb_dequant_req));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(capi.ListDesc), C.sizeof(capi.MeshDesc), C.sizeof(capi.ListStreams), C.sizeof(capi.Streams),
                     C.sizeof(capi.BatchStreams), C.sizeof(capi.QuantReq), C.sizeof(capi.DequantReq)]


@pytest.mark.skipif(not os.path.exists(capi.LIB_PATH), reason="libharry_b200.so not built")
def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.HarryError, match="no CUDA device|CPU fallback"):
        capi.Context(0)
