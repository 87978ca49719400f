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


def test_host_side_fallbacks_of_the_unet_and_decoder_ops_on_cpu():
    """ops.conv1x1 / ops.timeconv_decoder / conv3x3_valid route to the CUDA kernels only for CUDA tensors; on CPU tensors
    they must reproduce the plain torch modules exactly (this is what the CPU port and the gloo tests run through)."""
    import torch
    from neural_pde_surrogates_b200 import ops
    torch.manual_seed(0)
    conv = torch.nn.Conv2d(12, 9, 1)
    x = torch.randn(2, 12, 5, 7)
    assert torch.equal(ops.conv1x1(x, conv), conv(x))
    c3 = torch.nn.Conv2d(4, 6, 3)
    x3 = torch.randn(2, 4, 9, 8, requires_grad=True)
    assert torch.equal(ops.conv3x3_valid(x3, c3), c3(x3))
    tw, nc = 25, 1
    dec = torch.nn.Sequential(torch.nn.Conv1d(nc, 2 * nc, 13, stride=2), torch.nn.GELU(), torch.nn.Conv1d(2 * nc, nc, 8))
    z = torch.randn(2, 3 * tw * nc, 4, 3)
    ref = dec(z.permute(0, 2, 3, 1).reshape(2 * 4 * 3, nc, 3 * tw)).view(2, 4, 3, nc, tw).permute(0, 3, 4, 1, 2)
    assert torch.equal(ops.timeconv_decoder(z, dec, nc, tw), ref)
    assert ops._ranges(385, 208) == [(0, 193), (193, 192)] and ops._ranges(192, 208) == [(0, 192)]
    assert sum(n for _, n in ops._ranges(1000, 255)) == 1000 and max(n for _, n in ops._ranges(1000, 255)) <= 255


def test_dgrad_as_forward_conv_matches_autograd_on_cpu():
    """The input gradient of a stride-1 valid conv computed as a forward conv of the padded output gradient with the
    flipped, transposed filter (ops.ConvValidDgradAsForwardFunction) equals autograd's, in float64 on the host."""
    import torch
    from neural_pde_surrogates_b200 import ops
    torch.manual_seed(1)
    for k in (3, 2):
        w = torch.randn(5, 4, k, k, dtype=torch.float64, requires_grad=True)
        b = torch.randn(5, dtype=torch.float64, requires_grad=True)
        x = torch.randn(2, 4, 8, 7, dtype=torch.float64, requires_grad=True)
        y = ops.ConvValidDgradAsForwardFunction.apply(x, w, b)
        g = torch.randn_like(y)
        gx, gw, gb = torch.autograd.grad(y, (x, w, b), g)
        rx, rw, rb = torch.autograd.grad(torch.nn.functional.conv2d(x, w, b), (x, w, b), g)
        assert torch.allclose(gx, rx, atol=1e-12) and torch.allclose(gw, rw, atol=1e-12) and torch.allclose(gb, rb, atol=1e-12)
