"""B200-native mirror of the reference's `models/enc_proc_dec_components/proc_fno.py`.

Same class names, constructor signatures, attribute / parameter names (=> identical `state_dict` keys, shapes and
dtypes) and the same assertions as the reference, but `forward` runs the hand-written sm_100a kernel chain of
`ops.fno_block` instead of torch.fft + einsum + Conv2d + GELU:

    reference                                            here
    SpectralConv2d.forward   proc_fno.py:257-288         K1 -> K2 -> K3a -> K3b (no 1x1 term)
    FNO_Layer.forward        proc_fno.py:133-155         K1 -> K2 -> K3a -> K3b (+1x1, bias, activation fused)
    FNO.forward              proc_fno.py:73-83           the layer chain, reading h and the conditioning channels
                                                         through two pointers instead of torch.cat

Only the 2-D operator is native (the twophase configs); SpectralConv1d/3d and FiLM conditioning are listed as
"next" in SURVEY.md §8(f) and raise NotImplementedError instead of silently running something else.
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from .interfaces import D, M


class SpectralConv2d(nn.Module):
    """2-D Fourier layer: pruned rfft2, per-mode complex channel mixing, pruned irfft2 (proc_fno.py:225-288)."""

    def __init__(self, in_channels, out_channels, modes: tuple, feature_transform=False, feature_transform_dim=6,
                 transform_mode=1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.modes1 = modes[0]
        self.modes2 = modes[1]
        self.scale = 1 / (in_channels * out_channels)
        # same init, same order of RNG draws as the reference (proc_fno.py:239-243)
        self.weights1 = nn.Parameter(
            self.scale * torch.rand(in_channels, out_channels, self.modes1, self.modes2, dtype=torch.cfloat))
        self.weights2 = nn.Parameter(
            self.scale * torch.rand(in_channels, out_channels, self.modes1, self.modes2, dtype=torch.cfloat))
        self.feature_transform = feature_transform
        self.feature_transform_dim = feature_transform_dim
        self.transform_mode = transform_mode
        if feature_transform:
            raise NotImplementedError("FiLM conditioning of the spectral weights (proc_fno.py:271-284) is not part of "
                                      "the B200 hot path; use cond_mode='concat' as all twophase configs do")

    def forward(self, x, p=None):
        return ops.fno_block(x, None, None, self.weights1, self.weights2, None, None, ops.ACT_NONE)


def get_spectral_conv_with_right_spatial_dim(spatial_dim, **kwargs):
    if spatial_dim == 2:
        return SpectralConv2d(**kwargs)
    if spatial_dim in (1, 3):
        raise NotImplementedError(f"SpectralConv{spatial_dim}d is outside the B200 hot path (2-D twophase configs only)")
    raise NotImplementedError(f'only 0<x<=3d convs implemented so far, but found spatial dim {spatial_dim}!')


def _conv_nd(spatial_dim, **kwargs):
    if spatial_dim == 2:
        return nn.Conv2d(**kwargs)
    raise NotImplementedError(f"only the 2-D path is native; found spatial dim {spatial_dim}")


class FNO_Layer(nn.Module):
    """act(SpectralConv(x) + w(x) [+ w2(x)])   (proc_fno.py:87-155)."""

    def __init__(self, hidden_dim, num_spatial_dims: int = 1, kernel_size=1, modes=16, activation=nn.GELU,
                 activation_params=None, feature_transform=False, feature_transform_dim=6, transform_mode=0,
                 hidden_dim_out=None, conv_mode="single", padding_mode="circular"):
        super().__init__()
        self.num_spatial_dims = num_spatial_dims
        assert conv_mode in ["single", "double"]
        self.conv_mode = conv_mode
        if isinstance(modes, int):
            modes = tuple([modes for _ in range(num_spatial_dims)])
        assert len(modes) == num_spatial_dims, 'modes should be int or tuple of ints with length equal to spatial dim!'
        self.modes = modes
        if hidden_dim_out is None:
            hidden_dim_out = hidden_dim
        self.conv = get_spectral_conv_with_right_spatial_dim(
            spatial_dim=num_spatial_dims, in_channels=hidden_dim, out_channels=hidden_dim_out, modes=modes,
            feature_transform=feature_transform, feature_transform_dim=feature_transform_dim,
            transform_mode=transform_mode)
        if conv_mode == "single":
            self.w = _conv_nd(num_spatial_dims, in_channels=hidden_dim, out_channels=hidden_dim_out,
                              kernel_size=kernel_size, padding='same', padding_mode=padding_mode)
        else:
            self.w = _conv_nd(num_spatial_dims, in_channels=hidden_dim, out_channels=hidden_dim_out, kernel_size=1,
                              padding='same')
            self.w2 = _conv_nd(num_spatial_dims, in_channels=hidden_dim, out_channels=hidden_dim_out,
                               kernel_size=kernel_size, padding='same', padding_mode=padding_mode)
        if activation is None:
            self.act = None
        else:
            self.act = activation(**(activation_params or {}))

    # ---- fused entry used by FNO / UFNO: h and the conditioning channels arrive separately (no torch.cat),
    #      `res` is the U-Net branch and `act` the block activation (proc_ufno.py:111-118)
    def fused(self, h, vb=None, res=None, act=None):
        spat = h.shape[-self.num_spatial_dims:]
        for i, s in enumerate(spat):
            if i == len(spat) - 1:
                assert self.modes[i] <= s // 2 + 1, \
                    'modes should be at most the spatial dim // 2 + 1 for the last spatial dimension!'
            else:
                assert self.modes[i] <= s, 'modes should be at most the spatial dim all but the last spatial dimensions!'
        w_is_1x1 = tuple(self.w.kernel_size) == (1,) * self.num_spatial_dims
        code = ops.act_code(act)
        if w_is_1x1:
            extra = res
            if self.conv_mode == "double":
                x = h if vb is None else torch.cat([h, vb], dim=1)
                extra = self.w2(x) if res is None else self.w2(x) + res
            y = ops.fno_block(h, vb, extra, self.conv.weights1, self.conv.weights2, self.w.weight, self.w.bias,
                              ops.ACT_NONE if code is None else code)
        else:
            # k>1 local conv stays on torch/cuDNN; the spectral term, residual and activation stay fused
            x = h if vb is None else torch.cat([h, vb], dim=1)
            extra = self.w(x)
            if self.conv_mode == "double":
                extra = extra + self.w2(x)
            if res is not None:
                extra = extra + res
            y = ops.fno_block(h, vb, extra, self.conv.weights1, self.conv.weights2, None, None,
                              ops.ACT_NONE if code is None else code)
        return act(y) if code is None else y

    def forward(self, x, p=None):
        return self.fused(x, None, None, self.act)


class FNO(nn.Module):
    """Pure FNO processor (proc_fno.py:22-83)."""
    model_interface = M.AR_TB
    data_interface = [D.sim1d, D.sim1d_var_t, D.sim2d]

    def __init__(self, pde, num_spatial_dims: int = 1, n_cond: int = 0, hidden_features: int = 128,
                 fno_modes: int = 48, hidden_blocks: int = 4, cond_mode: str = "concat", fno_kernel_size: int = 1,
                 fno_conv_mode: str = "single", padding_mode: str = "circular", **kwargs):
        super().__init__()
        self.pde = pde
        self.num_spatial_dims = num_spatial_dims
        self.cond_mode = cond_mode
        assert self.cond_mode in ["film", "concat", None], "Incorrect conditioning mode supplied"
        if self.cond_mode == "film":
            feature_transform, feature_transform_dim, hidden_dim_in = n_cond > 0, n_cond, hidden_features
        elif self.cond_mode == "concat":
            feature_transform, feature_transform_dim, hidden_dim_in = False, 0, hidden_features + n_cond
        else:
            feature_transform, feature_transform_dim, hidden_dim_in = False, 0, hidden_features
        self.fno_layers = nn.ModuleList([FNO_Layer(
            hidden_dim=hidden_dim_in, hidden_dim_out=hidden_features, num_spatial_dims=num_spatial_dims,
            modes=fno_modes, feature_transform=feature_transform, feature_transform_dim=feature_transform_dim,
            kernel_size=fno_kernel_size, conv_mode=fno_conv_mode,
            padding_mode=padding_mode if padding_mode != "ones" else "zeros",
        ) for _ in range(hidden_blocks)])

    def __repr__(self):
        return f'FNO{self.num_spatial_dims}D'

    def forward(self, h, variables=None, variables_broadcast=None, pos=None):
        for layer in self.fno_layers:
            if self.cond_mode == "film":
                h = layer(h, p=variables)
            elif self.cond_mode == "concat":
                h = layer.fused(h, variables_broadcast, None, layer.act)
        return h
