"""CPU suite: the C-ABI library loads and exports every symbol include/pdes_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from neural_pde_surrogates_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "pdes_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pdes_[a-z0-9_]+)\s*\(", text)))


def test_binding_covers_header():
    assert sorted(_native.EXPORTED_SYMBOLS) == header_symbols()


def test_cuda_library_exports_every_symbol():
    if not os.path.exists(_native.LIB_PATH):
        import subprocess
        subprocess.check_call([os.path.join(ROOT, "build.sh")])
    lib = ctypes.CDLL(_native.LIB_PATH)
    for sym in header_symbols():
        assert hasattr(lib, sym), sym
    _native.bind(lib)
    assert lib.pdes_is_cuda_build() == 1 and lib.pdes_version() >= 100
    # host-only helpers are safe to call without a GPU
    assert lib.pdes_tables_floats(96, 64, 10, 10) > 0
    assert lib.pdes_block_fwd_workspace_floats(4, 193, 192, 96, 64, 10, 10) > 0
    assert lib.pdes_mix_suggest_splits(4, 193, 192, 10, 10) >= 1


def test_tables_match_oracle_constants():
    import numpy as np
    from oracle import spectral_oracle as so
    lib = _native.bind(ctypes.CDLL(_native.LIB_PATH))
    H, W, m1, m2 = 12, 8, 3, 5
    n = lib.pdes_tables_floats(H, W, m1, m2)
    buf = np.zeros(n, dtype=np.float32)
    assert lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data) == 0
    herm = buf[-8:][:m2]
    assert np.allclose(herm, so.hermitian_scale(H, W, m2), rtol=1e-7)
    assert lib.pdes_tables_fill(H, W, 13, m2, buf.ctypes.data) == _native.PDES_ERR_ARG
    with pytest.raises(ValueError):
        _native.check(lib, lib.pdes_tables_fill(H, W, m1, 6, buf.ctypes.data))
