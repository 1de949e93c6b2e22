"""CPU checks of the drop-in boundary: the shared library loads and exports every symbol include/mkhe.h
declares; without a GPU the product fails loudly instead of falling back to anything."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mkhe.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mkhe_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("mkhe_decompose", "mkhe_external_product_hoisted", "mkhe_mul_relin_hoisted", "mkhe_rotate_hoisted",
                 "mkhe_rescale", "mkhe_ckks_mul_relin", "mkhe_bfv_mul_relin_hoisted", "mkhe_bfv_quantize"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build_cuda()
    dll = ctypes.CDLL(g.LIB)
    missing = [s for s in declared_symbols() if not hasattr(dll, s)]
    assert not missing, missing
    dll.mkhe_version.restype = ctypes.c_char_p
    assert b"sm_100a" in dll.mkhe_version()


def test_no_cpu_fallback_without_gpu():
    from mkhe_kklss_b200 import _lib, params as PR
    lib = _lib.default_library()
    if lib.device_count() > 0:
        pytest.skip("a GPU is present")
    lit = PR.CKKS_PN14QP439.at_logn(12)
    with pytest.raises(_lib.MkheError):
        _lib.Context(lit.logN, lit.Q, lit.P)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mkhe_kklss_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/", "").lower() or f == "params.py" or "import" not in text, f
                assert "from oracle" not in text and "import oracle" not in text and "mkhe_oracle" not in text, f
                assert "libmkhe_emu" not in text, f
