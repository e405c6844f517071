"""CPU-side checks of the boundary: the C-ABI library builds for sm_100a, loads, and exports every
symbol include/graft_fem.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re

from dealii_adapter_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "graft_fem.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gf_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(capi.EXPORTED_SYMBOLS)


def test_library_builds_loads_and_exports_all_symbols():
    path = build.build_cuda()
    lib = ctypes.CDLL(path)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_create_fails_loudly_without_a_device_or_with_bad_input():
    import numpy as np
    L = capi.lib()
    d = capi.GfDesc()
    h = ctypes.c_void_p()
    d.dim = 5
    rc = L.gf_create(ctypes.byref(d), ctypes.byref(h))
    assert rc != capi.GF_OK and not h.value
    assert b"dim" in L.gf_last_error(None)


def test_product_does_not_reference_the_oracle():
    pkg = os.path.join(ROOT, "dealii_adapter_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(base, f), errors="ignore").read()
                assert "oracle_py" not in text and "liboracle" not in text.replace("LIB_ORACLE", "") \
                    or f == "build.py", f
