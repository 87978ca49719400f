"""B200-native mirror of the reference's `models/enc_proc_dec_components/proc_ufno.py` (UFNO processor).

Per block the reference runs (proc_ufno.py:105-119)
    h_in   = cat([h, vb])                # materialised concat
    h_fno  = FNO_Layer(h_in)             # rfft2, 2x einsum, zero-pad, irfft2, 1x1 conv, add      (~15 launches)
    h_unet = UNetModern(h, vb)
    h      = GELU(h_fno + h_unet)        # two more element-wise passes
Here `FNO_Layer.fused` does everything except the U-Net in one kernel chain and its last kernel (K3b) consumes
`h_unet` as the residual operand and applies the GELU while the accumulator tile is still in registers.
"""
from __future__ import annotations

from typing import List, Tuple, Union

from torch import nn

from .interfaces import D, M
from .proc_fno import FNO_Layer
from .unet_branch import UNetModern


class UFNO(nn.Module):
    model_interface = M.AR_TB
    data_interface = [D.sim1d, D.sim1d_var_t, D.sim2d]

    def __init__(self, pde, num_spatial_dims: int = 1, n_cond: int = 0, hidden_features: int = 128,
                 hidden_blocks: int = 4, cond_mode: str = "concat", padding_mode: str = "circular",
                 # FNO specific
                 fno_modes: int = 48, fno_kernel_size: int = 1, fno_conv_mode: str = "single",
                 # UNet specific
                 activation: nn.Module = nn.GELU(), norm: bool = False,
                 ch_mults: Union[Tuple[int, ...], List[int]] = (1, 1, 1),
                 is_attn: Union[Tuple[bool, ...], List[bool]] = (False, False, False),
                 mid_attn: bool = False, n_blocks: int = 1, use1x1: bool = True, **kwargs):
        super().__init__()
        self.pde = pde
        self.num_spatial_dims = num_spatial_dims
        self.cond_mode = cond_mode
        self.activation = activation
        assert self.cond_mode in ["film", "concat", None], "Incorrect conditioning mode supplied"
        if self.cond_mode == "film":
            feature_transform, feature_transform_dim, hidden_dim_in = n_cond > 0, n_cond, hidden_features
        elif self.cond_mode == "concat":
            feature_transform, feature_transform_dim, hidden_dim_in = False, 0, hidden_features + n_cond
        else:
            feature_transform, feature_transform_dim, hidden_dim_in = False, 0, hidden_features
        # construction order == reference (all FNO layers first, then all U-Nets) so a seeded init draws the
        # same random numbers for the same parameters
        self.fno_layers = nn.ModuleList([FNO_Layer(
            hidden_dim=hidden_dim_in, hidden_dim_out=hidden_features, num_spatial_dims=num_spatial_dims,
            modes=fno_modes, feature_transform=feature_transform, feature_transform_dim=feature_transform_dim,
            kernel_size=fno_kernel_size, conv_mode=fno_conv_mode,
            padding_mode=padding_mode if padding_mode != "ones" else "zeros", activation=None,
        ) for _ in range(hidden_blocks)])
        self.unet_layers = nn.ModuleList([UNetModern(
            pde=pde, num_spatial_dims=num_spatial_dims, n_cond=n_cond, hidden_features=hidden_features,
            cond_mode=cond_mode, activation=activation, norm=norm, ch_mults=ch_mults, is_attn=is_attn,
            mid_attn=mid_attn, n_blocks=n_blocks, use1x1=use1x1, padding_mode=padding_mode,
        ) for _ in range(hidden_blocks)])

    def __repr__(self):
        return f'U-FNO{self.num_spatial_dims}D'

    def forward(self, h, variables=None, variables_broadcast=None, pos=None):
        for fno_layer, unet in zip(self.fno_layers, self.unet_layers):
            if self.cond_mode not in ("film", "concat"):
                raise ValueError(f"Unknown cond_mode {self.cond_mode}")
            h_unet = unet(h=h, variables_broadcast=variables_broadcast, pos=pos)
            if self.cond_mode == "film":
                h = fno_layer.fused(h, None, h_unet, self.activation)
            else:
                h = fno_layer.fused(h, variables_broadcast, h_unet, self.activation)
        return h
