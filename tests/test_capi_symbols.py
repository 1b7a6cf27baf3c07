"""CPU: the C-ABI library loads and exports every symbol include/ubgl.h
declares (no compute calls without a GPU), and the host side fails loudly
instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ubgl.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ubgl_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    import ubootgl_b200
    L = ctypes.CDLL(ubootgl_b200.capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ubgl.h but not exported"
    # and the ctypes binding covers the whole header
    assert set(names) == set(ubootgl_b200.capi.SYMBOLS)


def test_version_and_no_cpu_fallback():
    import numpy as np
    import ubootgl_b200 as u
    assert u.lib.ubgl_version() == 100
    if u.lib.ubgl_device_count() == 0:
        with pytest.raises(u.UbglError):
            u.Simulation(np.ones((16, 16), np.float32))
        with pytest.raises(u.UbglError):
            u.MG(64, 64)
        with pytest.raises(u.UbglError):
            u.capi.rbgs(np.zeros((16, 16), np.float32), np.zeros((16, 16), np.float32),
                        np.ones((16, 16), np.float32), 0.1)
        assert b"CUDA" in u.lib.ubgl_last_error() or b"device" in u.lib.ubgl_last_error()


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (checked statically)."""
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "ubootgl_b200")):
        if "_build" in dp or "_lib" in dp:
            continue
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"(import|from)\s+oracle|#include\s+[\"<].*oracle|liboracle|libubgl_ref", txt):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
