"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU and exports
exactly the symbols include/diffqc_b200.h declares (no compute calls here)."""
import os
import re

import pytest

import diffquantum_b200 as dq
from diffquantum_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "diffqc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dq_[A-Za-z0-9_]+)\s*\(", txt)))


def test_library_is_built():
    assert os.path.isfile(_lib.LIB_PATH), "run __graft_entry__.build() first"


def test_header_and_binding_agree():
    syms = header_symbols()
    assert syms, "no symbols parsed from the header"
    assert sorted(_lib.SIGNATURES) == syms


def test_library_exports_every_declared_symbol():
    lib = dq.load()
    for s in header_symbols():
        assert hasattr(lib, s), "missing export %s" % s
    assert lib.dq_version() == b"dev"            # diffqc.__version__ (diffqc.cc:227)


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(Exception) as ei:
        dq.Context(0)
    assert "diffqc_b200" in str(ei.value)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "diffquantum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src, "%s mentions the oracle" % f
