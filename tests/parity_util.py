"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_l2(a, b):
    a = a.detach().cpu() if isinstance(a, torch.Tensor) else torch.as_tensor(np.asarray(a))
    b = b.detach().cpu() if isinstance(b, torch.Tensor) else torch.as_tensor(np.asarray(b))
    if a.is_complex():
        a = torch.view_as_real(a)
    if b.is_complex():
        b = torch.view_as_real(b)
    a, b = a.double(), b.double()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def load_prefixed_state(module, npz, prefix):
    sd = {k[len(prefix):]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(prefix)}
    module.load_state_dict(sd)
    return module


def tiny_model(device="cpu"):
    """The model of tests/golden/model_tiny.npz (cfg_twophase_ufno shrunk: 24x16 grid, width 16, modes 4, 2 blocks)."""
    import neural_pde_surrogates_b200 as npb
    g = golden("model_tiny.npz")
    pde = npb.TwoPhasePDE(24, 16)
    model = npb.build_twophase_model(pde=pde, hidden_features=16, fno_modes=4, hidden_blocks=2)
    load_prefixed_state(model, g, "sd_")
    return model.to(device), pde, g


def grads_close(named_params, ref_grads, tol, prefix=""):
    """Per-parameter relative L2 with a floor: gradients that are analytically zero (e.g. a bias in front of a
    per-channel GroupNorm) are pure rounding noise in both implementations, so the error is measured against
    max(|ref|, 1e-3 * largest gradient norm in the model)."""
    items = [(k, p.grad) for k, p in named_params]
    floor = 3e-3 * max(rel_norm(ref_grads(k)) for k, _ in items)
    worst = ("", 0.0)
    for k, gr in items:
        r = ref_grads(k)
        a = torch.view_as_real(gr).detach().cpu().double() if gr.is_complex() else gr.detach().cpu().double()
        b = torch.as_tensor(r)
        b = torch.view_as_real(b).double() if b.is_complex() else b.double()
        err = (a - b).norm().item() / max(b.norm().item(), floor)
        if err > worst[1]:
            worst = (k, err)
    assert worst[1] < tol, f"{prefix}{worst[0]}: rel L2 {worst[1]:.3e}"
    return worst


def rel_norm(r):
    b = torch.as_tensor(r)
    b = torch.view_as_real(b) if b.is_complex() else b
    return b.double().norm().item()
