"""libbnuts.so (the CUDA engine) loads on a CPU-only host and exports every symbol
include/bnuts.h declares.  No compute call is made here."""
import ctypes
import os
import re

from conftest import ROOT, CUDA_SO


def _declared():
    h = open(os.path.join(ROOT, "include", "bnuts.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(bnuts_[a-z_0-9]+)\s*\(", h)))


def test_header_and_binding_agree(bn):
    assert sorted(bn.EXPORTS) == _declared()


def test_cuda_library_exports_every_symbol():
    assert os.path.exists(CUDA_SO), "build the CUDA engine first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(CUDA_SO)
    for name in _declared():
        assert hasattr(lib, name), name


def test_product_never_loads_the_oracle():
    pkg = os.path.join(ROOT, "inplacedhmc.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".h", ".cu", ".cpp", ".jl")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "libbnuts_oracle" not in txt and "oracle/" not in txt.replace("the oracle/", ""), f
