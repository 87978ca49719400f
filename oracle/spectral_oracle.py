"""CPU oracle for the U-FNO / FNO spectral block (TEST INFRASTRUCTURE ONLY).

This file is a float64 numpy restatement of the reference's spectral hot path.  It is the
checker for the CUDA kernels; it is never imported by the product package
(`neural_pde_surrogates_b200`).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import it.

Parity pin: the reference ships no tests or golden vectors, so this oracle
is pinned by (i) `tests/golden/*.npz`, produced by *executing the reference modules* in the
build container (`tests/golden/make_golden.py`; checked by `tests/test_oracle_golden.py`), and (ii)
`tests/test_oracle_vs_reference.py`, which imports the reference live (from `/root/reference/src` or the verbatim copy
in `oracle/_ref/src`, see `oracle/reference_loader.py`) and compares on freshly drawn inputs.

Reference lines restated here (paths relative to /root/reference/src):
  * models/enc_proc_dec_components/proc_fno.py:257-288  SpectralConv2d.forward
      rfft2 -> keep rows [0,m1) and [H-m1,H), cols [0,m2) -> per-mode channel mix -> irfft2
  * proc_fno.py:253-255  compl_mul2d  einsum("bixy,ioxy->boxy")
  * proc_fno.py:133-155  FNO_Layer.forward   act(conv(x) + w(x))
  * models/enc_proc_dec_components/proc_ufno.py:105-119  UFNO.forward
      h = activation(FNO_Layer(cat[h, vb]) + UNetModern(h, vb))

The restatement uses *dense truncated DFT matrices* (no FFT), i.e. exactly the contraction
structure of the CUDA kernels:
    X[b,i,k,l] = sum_{h,w} x[b,i,h,w] * exp(-2 pi i (kx_k h / H + l w / W))          (K1)
    O[b,o,k,l] = live_k * sum_i X[b,i,k,l] * Wt[i,o,k,l]                             (K2)
    y[b,o,h,w] = Re sum_{k,l} s_l * exp(+2 pi i (kx_k h / H + l w / W)) * O[b,o,k,l]  (K3)
with kx_k = k for k < m1 and H - 2 m1 + k for k >= m1, Wt = cat(weights1, weights2) along k,
s_l = c_l / (H W), c_0 = 1, c_l = 2 for 0 < l < W/2 and c_{W/2} = 1 (W even), and
live_k = 0 for first-block rows that the reference overwrites with the second block when
2 m1 > H (proc_fno.py:266-269 assigns the second slice after the first).
"""
from __future__ import annotations

import numpy as np
from scipy.special import erf

__all__ = [
    "kx_table", "live_rows", "hermitian_scale", "dft_fwd_pruned", "mode_mix", "mode_mix_dx", "mode_mix_dw",
    "inv_pruned", "spectral_conv2d_forward", "spectral_conv2d_backward", "gelu", "gelu_grad",
    "fno_block_forward", "fno_block_backward", "rel_l2",
]


def rel_l2(a, b) -> float:
    a = np.asarray(a); b = np.asarray(b)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))


def kx_table(H: int, m1: int) -> np.ndarray:
    """Frequency index of retained row k (proc_fno.py:266-269: rows [:m1] and [-m1:])."""
    k = np.arange(2 * m1)
    return np.where(k < m1, k, H - 2 * m1 + k)


def live_rows(H: int, m1: int) -> np.ndarray:
    """1 for rows whose value survives in out_ft; a first-block row with kx >= H-m1 is overwritten
    by the second block (proc_fno.py:268-269 runs after :266-267)."""
    k = np.arange(2 * m1)
    dead = (k < m1) & (k >= H - m1)
    return (~dead).astype(np.float64)


def hermitian_scale(H: int, W: int, m2: int) -> np.ndarray:
    """s_l = c_l/(H W): what irfft2 (proc_fno.py:287) applies to column l of a half spectrum."""
    l = np.arange(m2)
    c = np.full(m2, 2.0)
    c[l == 0] = 1.0
    if W % 2 == 0:
        c[l == W // 2] = 1.0
    return c / (H * W)


def _row_phase(H, m1, sign):
    kx = kx_table(H, m1)[:, None].astype(np.float64)
    h = np.arange(H)[None, :].astype(np.float64)
    return np.exp(sign * 2j * np.pi * kx * h / H)          # [2m1, H]


def _col_phase(W, m2, sign):
    l = np.arange(m2)[:, None].astype(np.float64)
    w = np.arange(W)[None, :].astype(np.float64)
    return np.exp(sign * 2j * np.pi * l * w / W)           # [m2, W]


def dft_fwd_pruned(x, m1: int, m2: int, lscale=None):
    """K1.  x [B,C,H,W] real -> X [B,C,2m1,m2] complex = rfft2(x)[:, :, rows, :m2] (proc_fno.py:261,267,269).
    `lscale[l]` optionally multiplies column l (used by the backward: GO = s_l * DFT(g))."""
    x = np.asarray(x, dtype=np.float64)
    H, W = x.shape[-2:]
    Eh = _row_phase(H, m1, -1.0)                           # [K,H]
    Ew = _col_phase(W, m2, -1.0)                           # [m2,W]
    Y = np.einsum("bchw,lw->bchl", x, Ew)
    X = np.einsum("kh,bchl->bckl", Eh, Y)
    if lscale is not None:
        X = X * np.asarray(lscale, dtype=np.float64)[None, None, None, :]
    return X


def _wt(w1, w2):
    return np.concatenate([np.asarray(w1, dtype=np.complex128), np.asarray(w2, dtype=np.complex128)], axis=2)


def mode_mix(X, w1, w2, H: int):
    """K2 (proc_fno.py:253-255, 266-269).  X [B,Cin,2m1,m2], w1/w2 [Cin,Cout,m1,m2] -> O [B,Cout,2m1,m2]."""
    m1 = w1.shape[2]
    O = np.einsum("bikl,iokl->bokl", X, _wt(w1, w2))
    return O * live_rows(H, m1)[None, None, :, None]


def mode_mix_dx(GO, w1, w2, H: int):
    """Adjoint of K2 w.r.t. X (torch convention dL/dRe + i dL/dIm): GX = sum_o GO * conj(Wt)."""
    m1 = w1.shape[2]
    GOm = GO * live_rows(H, m1)[None, None, :, None]
    return np.einsum("bokl,iokl->bikl", GOm, np.conj(_wt(w1, w2)))


def mode_mix_dw(X, GO, H: int, m1: int):
    """Adjoint of K2 w.r.t. the weights: GW[i,o,k,l] = sum_b conj(X) * GO; returns (gw1, gw2)."""
    GOm = GO * live_rows(H, m1)[None, None, :, None]
    GW = np.einsum("bikl,bokl->iokl", np.conj(X), GOm)
    return GW[:, :, :m1], GW[:, :, m1:]


def inv_pruned(O, H: int, W: int, lscale):
    """K3 spectral part.  O [B,C,2m1,m2] -> y [B,C,H,W] = Re sum_{k,l} lscale_l e^{+i..} O.
    With lscale = hermitian_scale this equals irfft2(zero-padded O, s=(H,W)) (proc_fno.py:265-269,287);
    with lscale = 1 it is the adjoint of `dft_fwd_pruned`."""
    O = np.asarray(O, dtype=np.complex128)
    m1 = O.shape[2] // 2
    m2 = O.shape[3]
    Eh = _row_phase(H, m1, +1.0)                           # [K,H]
    Ew = _col_phase(W, m2, +1.0) * np.asarray(lscale, dtype=np.float64)[:, None]   # [m2,W]
    Z = np.einsum("kh,bckl->bchl", Eh, O)
    return np.real(np.einsum("bchl,lw->bchw", Z, Ew))


def spectral_conv2d_forward(x, w1, w2):
    """SpectralConv2d.forward without FiLM (proc_fno.py:257-288). Returns (y, X)."""
    H, W = x.shape[-2:]
    m1, m2 = w1.shape[2], w1.shape[3]
    X = dft_fwd_pruned(x, m1, m2)
    O = mode_mix(X, w1, w2, H)
    return inv_pruned(O, H, W, hermitian_scale(H, W, m2)), X


def spectral_conv2d_backward(x, w1, w2, g, X=None):
    """Gradients of sum(g*y): returns (gx, gw1, gw2), complex grads in torch's convention."""
    H, W = x.shape[-2:]
    m1, m2 = w1.shape[2], w1.shape[3]
    if X is None:
        X = dft_fwd_pruned(x, m1, m2)
    GO = dft_fwd_pruned(g, m1, m2, lscale=hermitian_scale(H, W, m2))
    GX = mode_mix_dx(GO, w1, w2, H)
    gw1, gw2 = mode_mix_dw(X, GO, H, m1)
    gx = inv_pruned(GX, H, W, np.ones(m2))
    return gx, gw1, gw2


def gelu(x):
    """nn.GELU() default = exact erf form (proc_ufno.py:44, proc_fno.py:94)."""
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * x * (1.0 + erf(x / np.sqrt(2.0)))


def gelu_grad(x):
    x = np.asarray(x, dtype=np.float64)
    return 0.5 * (1.0 + erf(x / np.sqrt(2.0))) + x * np.exp(-0.5 * x * x) / np.sqrt(2.0 * np.pi)


def fno_block_forward(h, vb, w1, w2, wc, bias, res=None, act="gelu"):
    """One fused block tail.
        h_in = cat[h, vb]                                   proc_ufno.py:111 / proc_fno.py:78
        pre  = SpectralConv2d(h_in) + Conv2d_1x1(h_in)      proc_fno.py:142-146
        pre += res   (the U-Net branch, proc_ufno.py:117-118; None for the pure FNO layer)
        out  = act(pre)                                     proc_ufno.py:118 / proc_fno.py:153-154
    wc [Cout,Cin] (the [Cout,Cin,1,1] conv weight squeezed), bias [Cout] or None.  Returns (out, pre, X)."""
    h = np.asarray(h, dtype=np.float64)
    h_in = h if vb is None else np.concatenate([h, np.asarray(vb, dtype=np.float64)], axis=1)
    y, X = spectral_conv2d_forward(h_in, w1, w2)
    pre = y
    if wc is not None:
        pre = pre + np.einsum("oi,bihw->bohw", np.asarray(wc, dtype=np.float64), h_in)
    if bias is not None:
        pre = pre + np.asarray(bias, dtype=np.float64)[None, :, None, None]
    if res is not None:
        pre = pre + np.asarray(res, dtype=np.float64)
    out = gelu(pre) if act == "gelu" else pre
    return out, pre, X


def fno_block_backward(h, vb, w1, w2, wc, bias, res, act, g_out):
    """Gradients of sum(g_out*out).  Returns dict(dh, dvb, dres, dw1, dw2, dwc, dbias)."""
    h = np.asarray(h, dtype=np.float64)
    C0 = h.shape[1]
    h_in = h if vb is None else np.concatenate([h, np.asarray(vb, dtype=np.float64)], axis=1)
    _, pre, X = fno_block_forward(h, vb, w1, w2, wc, bias, res, act)
    g_pre = np.asarray(g_out, dtype=np.float64) * (gelu_grad(pre) if act == "gelu" else 1.0)
    gx, gw1, gw2 = spectral_conv2d_backward(h_in, w1, w2, g_pre, X)
    out = dict(dw1=gw1, dw2=gw2, dres=g_pre if res is not None else None)
    if wc is not None:
        wc = np.asarray(wc, dtype=np.float64)
        gx = gx + np.einsum("oi,bohw->bihw", wc, g_pre)
        out["dwc"] = np.einsum("bohw,bihw->oi", g_pre, h_in)
    if bias is not None:
        out["dbias"] = g_pre.sum(axis=(0, 2, 3))
    out["dh"] = gx[:, :C0]
    out["dvb"] = gx[:, C0:] if vb is not None else None
    return out
