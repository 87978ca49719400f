"""The U-Net branch of a U-FNO block, kept on PyTorch/cuDNN (SURVEY.md §8(f) "next #1").

Its output is the residual operand of the fused K3b epilogue.  This is a from-scratch restatement of the reference's
`UNetModern` (models/enc_proc_dec_components/proc_unet_modern.py:24-196) with the same submodule / parameter names,
so `state_dict`s are interchangeable, and the same arithmetic, including the reference's quirks:

  * `padding_mode="circular"` is passed to the 3x3 convs WITHOUT `padding=` (proc_unet_modern.py:78-81), so they
    are *valid* convolutions that shrink the map by 2; shapes are repaired by symmetric zero padding / cropping
    (`crop_Nd`, models/common.py:20-34), with the odd pixel going to the far side;
  * the up-sampling ConvTranspose2d(k=4, s=2) sees a 1-pixel circularly padded input (models/common.py:61-120);
  * conditioning channels are concatenated at every block and down-sampled by their own strided conv.
Only 2-D is implemented (the twophase configs).
"""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from .interfaces import D, M


def fit_to(t: torch.Tensor, spatial) -> torch.Tensor:
    """Zero-pad (or crop, for negative amounts) the trailing dims of `t` to `spatial`, split evenly with the odd
    pixel on the far side -- the behaviour of crop_Nd (models/common.py:20-34)."""
    pads = []
    for cur, want in zip(reversed(t.shape[-len(spatial):]), reversed(tuple(spatial))):
        lo = (want - cur) // 2
        pads += [lo, (want - cur) - lo]
    if not any(pads):
        return t
    return F.pad(t, pads)


def _conv_kwargs(padding_mode: str) -> dict:
    assert padding_mode in ["ones", "circular"]
    # "circular" without an explicit padding => padding 0 => valid convolution (reference quirk, see module doc)
    return dict(padding=1) if padding_mode == "ones" else dict(padding_mode="circular")


class ConvTranspose2d_padded(nn.ConvTranspose2d):
    """ConvTranspose2d applied to a circularly padded input (models/common.py:95-103)."""

    def __init__(self, pad, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.pad = pad

    def forward(self, x):
        p = self.pad
        x = torch.cat([x[..., -p:], x, x[..., :p]], dim=-1)
        x = torch.cat([x[..., -p:, :], x, x[..., :p, :]], dim=-2)
        return super().forward(x)


class ResidualBlock(nn.Module):
    """conv2(act(norm2(conv1(act(norm1(x)))))) + shortcut(x)   (proc_unet_modern.py:199-250)."""

    def __init__(self, in_channels, out_channels, activation=nn.GELU(), norm=False, n_groups=1, num_spatial_dims=2,
                 padding_kwargs=None):
        super().__init__()
        kw = padding_kwargs or {}
        self.activation = activation
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, **kw)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, **kw)
        self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else nn.Identity()
        self.norm1 = nn.GroupNorm(n_groups, in_channels) if norm else nn.Identity()
        self.norm2 = nn.GroupNorm(n_groups, out_channels) if norm else nn.Identity()

    def forward(self, x):
        y = ops.conv3x3_valid(ops.group_norm_act(x, self.norm1, self.activation), self.conv1)
        y = ops.conv3x3_valid(ops.group_norm_act(y, self.norm2, self.activation), self.conv2)
        skip = ops.conv1x1(x, self.shortcut) if isinstance(self.shortcut, nn.Conv2d) else x
        return fit_to(y, skip.shape[-2:]) + skip


class AttentionBlock(nn.Module):
    """Spatial self-attention as in the reference (proc_unet_modern.py:253-316); note its softmax runs over the
    *query* axis (dim=1 of the [b, i, j, h] scores)."""

    def __init__(self, in_channels, out_channels=None, n_heads=1, d_k=None, n_groups=1, num_spatial_dims=2):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        d_k = in_channels if d_k is None else d_k
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm = nn.GroupNorm(n_groups, in_channels)       # present in the state_dict, unused in forward
        self.projection = nn.Linear(in_channels, n_heads * d_k * 3)
        self.output = nn.Linear(n_heads * d_k, out_channels)
        self.scale = d_k ** -0.5
        self.n_heads, self.d_k = n_heads, d_k
        self.shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels else nn.Identity()

    def forward(self, x):
        b, _, *spatial = x.shape
        seq = x.view(b, self.in_channels, -1).permute(0, 2, 1)
        q, k, v = self.projection(seq).view(b, -1, self.n_heads, 3 * self.d_k).chunk(3, dim=-1)
        scores = torch.einsum("bihd,bjhd->bijh", q, k) * self.scale
        mixed = torch.einsum("bijh,bjhd->bihd", scores.softmax(dim=1), v).reshape(b, -1, self.n_heads * self.d_k)
        y = self.output(mixed) + self.shortcut(seq)
        return y.permute(0, 2, 1).reshape(b, self.out_channels, *spatial)


def _attn(flag, ch):
    return AttentionBlock(ch) if flag else nn.Identity()


class DownBlock(nn.Module):
    def __init__(self, in_channels, out_channels, has_attn=False, activation=nn.GELU(), norm=False,
                 num_spatial_dims=2, padding_kwargs=None):
        super().__init__()
        self.res = ResidualBlock(in_channels, out_channels, activation=activation, norm=norm, padding_kwargs=padding_kwargs)
        self.attn = _attn(has_attn, out_channels)

    def forward(self, x, variables_broadcast=None):
        if variables_broadcast is not None:
            x = torch.cat([x, variables_broadcast], dim=1)
        return self.attn(self.res(x)), variables_broadcast


class UpBlock(nn.Module):
    def __init__(self, in_channels, out_channels, has_attn=False, activation=nn.GELU(), norm=False,
                 num_spatial_dims=2, padding_kwargs=None):
        super().__init__()
        # input = current features + skip connection (+ conditioning): in_channels + out_channels
        self.res = ResidualBlock(in_channels + out_channels, out_channels, activation=activation, norm=norm,
                                 padding_kwargs=padding_kwargs)
        self.attn = _attn(has_attn, out_channels)

    def forward(self, x):
        return self.attn(self.res(x))


class MiddleBlock(nn.Module):
    def __init__(self, in_channels, out_channels, has_attn=False, activation=nn.GELU(), norm=False,
                 num_spatial_dims=2, padding_kwargs=None):
        super().__init__()
        self.res1 = ResidualBlock(in_channels, out_channels, activation=activation, norm=norm, padding_kwargs=padding_kwargs)
        self.attn = _attn(has_attn, out_channels)
        self.res2 = ResidualBlock(out_channels, out_channels, activation=activation, norm=norm, padding_kwargs=padding_kwargs)

    def forward(self, x, variables_broadcast=None):
        if variables_broadcast is not None:
            x = torch.cat([x, variables_broadcast], dim=1)
        return self.res2(self.attn(self.res1(x))), variables_broadcast


class Upsample(nn.Module):
    def __init__(self, n_channels, num_spatial_dims, padding_kwargs):
        super().__init__()
        if padding_kwargs.get("padding_mode") == "circular":
            self.conv = ConvTranspose2d_padded((4 - 1) // 2, n_channels, n_channels, kernel_size=4, stride=2)
        else:
            self.conv = nn.ConvTranspose2d(n_channels, n_channels, kernel_size=4, stride=2, **padding_kwargs)

    def forward(self, x):
        return self.conv(x)


class Downsample(nn.Module):
    def __init__(self, n_channels, num_spatial_dims, n_cond, padding_kwargs):
        super().__init__()
        self.conv = nn.Conv2d(n_channels, n_channels, kernel_size=3, stride=2, **padding_kwargs)
        if n_cond > 0:
            self.conv_variables_broadcast = nn.Conv2d(n_cond, n_cond, kernel_size=3, stride=2, **padding_kwargs)

    def forward(self, x, variables_broadcast=None):
        if variables_broadcast is not None:
            return self.conv(x), self.conv_variables_broadcast(variables_broadcast)
        return self.conv(x), None


class UNetModern(nn.Module):
    model_interface = M.AR_TB
    data_interface = [D.sim1d, D.sim2d, D.sim1d_var_t]

    def __init__(self, pde, num_spatial_dims: int = 1, n_cond: int = 0, hidden_features: int = 128,
                 cond_mode: str = "concat", activation: nn.Module = nn.GELU(), norm: bool = False,
                 ch_mults: Union[Tuple[int, ...], List[int]] = (1, 2, 2, 4),
                 is_attn: Union[Tuple[bool, ...], List[bool]] = (False, False, False, False),
                 mid_attn: bool = False, n_blocks: int = 2, use1x1: bool = False, padding_mode: str = "ones",
                 **kwargs) -> None:
        super().__init__()
        if num_spatial_dims != 2:
            raise NotImplementedError("the B200 build covers the 2-D twophase configs only")
        self.hidden_features = hidden_features
        self.num_spatial_dims = num_spatial_dims
        assert cond_mode in ["concat", None], "Incorrect conditioning mode supplied"
        self.cond_mode = cond_mode
        self.n_cond = 0 if cond_mode is None else n_cond
        pk = _conv_kwargs(padding_mode)
        self.activation = activation
        common = dict(activation=activation, norm=norm, padding_kwargs=pk)
        levels = len(ch_mults)

        down, ch = [], hidden_features
        for lvl in range(levels):
            wide = ch * ch_mults[lvl]
            for _ in range(n_blocks):
                down.append(DownBlock(ch + n_cond, wide, has_attn=is_attn[lvl], **common))
                ch = wide
            if lvl < levels - 1:
                down.append(Downsample(ch, num_spatial_dims, n_cond=n_cond, padding_kwargs=pk))
        self.down = nn.ModuleList(down)

        self.middle = MiddleBlock(ch + n_cond, ch, has_attn=mid_attn, **common)

        up = []
        for lvl in reversed(range(levels)):
            for _ in range(n_blocks):
                up.append(UpBlock(ch + n_cond, ch, has_attn=is_attn[lvl], **common))
            narrow = ch // ch_mults[lvl]
            up.append(UpBlock(ch + n_cond, narrow, has_attn=is_attn[lvl], **common))
            ch = narrow
            if lvl > 0:
                up.append(Upsample(ch, num_spatial_dims, padding_kwargs=pk))
        self.up = nn.ModuleList(up)

        self.norm = nn.GroupNorm(8, hidden_features) if norm else nn.Identity()
        if use1x1:
            self.final = nn.Conv2d(hidden_features, hidden_features, kernel_size=1)
        else:
            self.final = nn.Conv2d(hidden_features, hidden_features, kernel_size=3, **pk)

    def forward(self, h, variables_broadcast=None, pos=None):
        assert h.dim() == 2 + self.num_spatial_dims
        target = h.shape[-2:]
        skips, conds = [h], [variables_broadcast]
        vb = variables_broadcast
        for m in self.down:
            h, vb = m(h, vb)
            skips.append(h)
            conds.append(vb)
        h, vb = self.middle(h, vb)
        for m in self.up:
            if isinstance(m, Upsample):
                h = m(h)
                continue
            here = h.shape[-2:]
            parts = [h, fit_to(skips.pop(), here)]
            c = conds.pop()
            if c is not None:
                parts.append(fit_to(c, here))
            h = m(torch.cat(parts, dim=1))
        z = ops.group_norm_act(h, self.norm, self.activation)
        z = ops.conv1x1(z, self.final) if tuple(self.final.kernel_size) == (1, 1) else self.final(z)
        return fit_to(z, target)
