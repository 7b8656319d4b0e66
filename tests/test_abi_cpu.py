"""CPU-side checks of the C-ABI library: it loads without a GPU, exports every symbol include/gfs_b200.h declares,
fails loudly when asked to compute without a device, and the product package never imports the oracle."""
import os
import re

import pytest

from gridfluidsim3d_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "gfs_b200.h")).read()
    declared = set(re.findall(r"\b(gfs_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found in the header"
    lib = capi.load_library()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, "libgfs_b200.so lacks %s" % missing
    assert declared == set(capi.SYMBOLS), "capi.SYMBOLS and the header disagree: %s" % sorted(declared ^ set(capi.SYMBOLS))


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GfsError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gridfluidsim3d_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in text and "liboracle" not in text and "orc_" not in text, f
                assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), f


def test_error_convention():
    lib = capi.load_library()
    import ctypes as C
    err = C.c_int(7)
    k0, k1 = C.c_int(), C.c_int()
    lib.gfs_slab_range(16, 0, 0, C.byref(k0), C.byref(k1), C.byref(err))
    assert err.value == 0 and b"bad arguments" in lib.gfs_get_error_message()
    lib.gfs_slab_range(16, 4, 1, C.byref(k0), C.byref(k1), C.byref(err))
    assert err.value == 1 and (k0.value, k1.value) == (4, 8)
