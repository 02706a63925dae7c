"""The C-ABI shared library: loads, exports every symbol ``include/gwin_b200.h``
declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from gwin_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gwin_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gwin_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 12
    for n in names:
        assert getattr(lib, n) is not None, n
    assert sorted(_lib.SYMBOLS) == names
    assert b"sm_100a" in lib.gwin_version()


def test_library_is_built_for_sm100a_only():
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_argument_validation_and_no_cpu_fallback():
    lib = _lib.load()
    h = ctypes.c_void_p()
    resp = np.zeros(9)
    loc = np.zeros(3)
    data = np.zeros(2 * 64)
    dp = ctypes.POINTER(ctypes.c_double)

    def create(n_det=1, n_f=64, df=1.0):
        return lib.gwin_plan_create(ctypes.byref(h), 0, n_det, n_f, df, 0.0, 4.0, 0.0, 4.0,
                                    resp.ctypes.data_as(dp), loc.ctypes.data_as(dp),
                                    data.ctypes.data_as(dp), None, 0.0)
    assert create(n_det=0) == -1 and b"n_det" in lib.gwin_last_error()
    assert create(n_det=9) == -1
    assert create(df=0.0) == -1
    assert create(n_f=2) == -1
    import torch
    if not torch.cuda.is_available():
        # valid arguments but no device: must fail loudly, never compute on the host
        assert create() == -4
        assert b"no CPU fallback" in lib.gwin_last_error()
        with pytest.raises(_lib.GwinCudaError):
            _lib.Plan(0, 1.0, 0.0, 4.0, None, 4.0, resp, loc, np.zeros((1, 64), complex), None, None)
    # NULL plan handles are rejected, not dereferenced
    assert lib.gwin_plan_lognl(None, None, None) == -1
    assert lib.gwin_loglr_batch(None, 0, 1, None, None, None, None, None, None) == -1
    assert lib.gwin_plan_last_launches(None) == 0
    lib.gwin_plan_destroy(None)


def test_product_never_imports_oracle():
    """The product path may not route through the oracle (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "gwin_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "gwin_oracle" not in src and "c_oracle" not in src and \
                    "loglr_ref" not in src, f
