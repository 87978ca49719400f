"""CPU (PyTorch) port of the spectral block, differentiable -- TEST INFRASTRUCTURE ONLY.

`fno_block_torch` restates, with library ops, what the CUDA chain computes; `cpu_port(model)` swaps it into a model
built from the product modules so that the *whole* model runs on the host.  Used by: the CPU parity tests (vs the
reference and vs the golden fixtures), the gloo data-parallel tests, `smoke()`'s checker and the `cpu_baseline` /
`--impl reference` legs of bench.py.  The product package never imports this file.

Follows reference src/models/enc_proc_dec_components/proc_fno.py:257-288 (rfft2 -> two mode blocks ->
einsum("bixy,ioxy->boxy") -> irfft2), :142-155 (+ 1x1 conv, activation) and proc_ufno.py:111-118.
"""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F


def spectral_conv2d_torch(x, w1, w2):
    B, _, H, W = x.shape
    _, Cout, m1, m2 = w1.shape
    xf = torch.fft.rfft2(x)
    yf = torch.zeros(B, Cout, H, W // 2 + 1, dtype=torch.cfloat, device=x.device)
    yf[:, :, :m1, :m2] = torch.einsum("bixy,ioxy->boxy", xf[:, :, :m1, :m2], w1)
    yf[:, :, -m1:, :m2] = torch.einsum("bixy,ioxy->boxy", xf[:, :, -m1:, :m2], w2)
    return torch.fft.irfft2(yf, s=(H, W))


def fno_block_torch(h, vb, res, w1, w2, wc, bias, act):
    """Same signature as neural_pde_surrogates_b200.ops.fno_block (act: 0 none, 1 exact GELU)."""
    x = h if vb is None else torch.cat([h, vb], dim=1)
    y = spectral_conv2d_torch(x, w1, w2)
    if wc is not None:
        y = y + F.conv2d(x, wc.reshape(wc.shape[0], wc.shape[1], 1, 1), bias)
    if res is not None:
        y = y + res
    return F.gelu(y) if act == 1 else y


@contextlib.contextmanager
def cpu_port():
    """Within this context the product modules call the torch port instead of the CUDA chain."""
    from neural_pde_surrogates_b200 import ops
    saved = ops.fno_block
    ops.fno_block = fno_block_torch
    try:
        yield
    finally:
        ops.fno_block = saved
