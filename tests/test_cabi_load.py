"""The C-ABI library builds, loads and exports every symbol include/lhvi.h declares.
No compute call is made here (no GPU in this tier)."""
import ctypes
import os
import re

import pytest

import lhvi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "lhvi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lhvi_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported():
    from lhvi_b200 import _cabi, build
    build.build()
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert set(names) == set(_cabi.SYMBOLS)
    for name in names:
        assert hasattr(lib, name), name


def test_binding_loads_and_reports_errors():
    from lhvi_b200 import _cabi
    lib = _cabi.load(build_if_missing=True)
    assert lib.lhvi_abi_version() == _cabi.ABI_VERSION
    # argument validation happens before any CUDA call, so it is checkable without a GPU
    rc = lib.lhvi_factor_expect_grad(None, None, 0, 0, None)
    assert rc == -1 and b"null" in lib.lhvi_last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, lib)
    m = _cabi.LhviModel()
    m.dtype, m.K, m.T = 7, 2, 3
    g = _cabi.LhviGroup()
    assert lib.lhvi_factor_expect_grad(ctypes.byref(m), ctypes.byref(g), 0, 0, None) == -1
    m.dtype, m.K = _cabi.LHVI_F64, 99
    assert lib.lhvi_factor_expect_grad(ctypes.byref(m), ctypes.byref(g), 0, 0, None) == -2
    assert lib.lhvi_step_tick(None, 0.9, 0.999, None) == -1


def test_struct_layout_matches_header():
    """sizeof the ctypes mirrors == sizeof the C structs (compiled with gcc from the header)."""
    import subprocess
    import tempfile
    from lhvi_b200 import _cabi
    src = '#include <stdio.h>\n#include "lhvi.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(lhvi_group), sizeof(lhvi_model), sizeof(lhvi_exchange), sizeof(lhvi_optim), sizeof(lhvi_h2), sizeof(lhvi_gabp));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(d, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sg, sm, sx, so, sh, sb = map(int, subprocess.check_output([exe]).split())
    assert sg == ctypes.sizeof(_cabi.LhviGroup)
    assert sm == ctypes.sizeof(_cabi.LhviModel)
    assert sx == ctypes.sizeof(_cabi.LhviExchange)
    assert so == ctypes.sizeof(_cabi.LhviOptim)
    assert sh == ctypes.sizeof(_cabi.LhviH2)
    assert sb == ctypes.sizeof(_cabi.LhviGabp)


def test_engine_refuses_to_run_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    syn = lhvi_b200.synthetic
    from lhvi_b200.engine import DeviceEngine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        DeviceEngine(syn.gaussian_grid(3, 1, 3))
