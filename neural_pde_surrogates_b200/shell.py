"""Per-step model shell around the processors: conditioning broadcast, point-wise encoder, time-conv decoder and
the output-constraint wrapper.  Kept on PyTorch (SURVEY.md §8 a-17 / (f) "next #2"); it exists here so that the
full `cfg_twophase_ufno` model can be built, trained, rolled out and benchmarked without the reference tree.

From-scratch restatement with the reference's parameter names (=> interchangeable state_dicts):
  encoder.encoder.{0,2}        ElementWise      models/enc_proc_dec_components/enc_grid.py:24-50
  processor.{i}.*              FNO / UFNO       proc_fno.py / proc_ufno.py
  decoder.pre_decoder, decoder.decoder.{0,2}    TimeConvDense  dec_grid.py:97-146 (+ add_delta :8-23)
  shell                        EncProcDec       models/enc_proc_dec.py:41-183
  output constraints           activation_wrapper   models/activation_wrapper.py:9-108
"""
from __future__ import annotations

import copy
import math
from types import SimpleNamespace

import torch
from torch import nn

from . import ops
from .proc_fno import FNO
from .proc_ufno import UFNO
from .unet_branch import UNetModern


class TwoPhasePDE(SimpleNamespace):
    """Metadata the models read from the dataset's PDE object (src/pdes/base.py:34-52)."""

    def __init__(self, H=96, W=64, nt=501, tmax=5.0, L1=1.5, L2=1.0, n_cond_static=0, n_cond_spatial=1, name="twophase"):
        super().__init__(tmin=0.0, tmax=tmax, nt=nt, name=name, n_cond_static=n_cond_static, n_cond_dynamic=0,
                         n_cond_spatial=n_cond_spatial, L1=L1, L2=L2, L=[L1, L2], nx1=H, nx2=W, dt=tmax / (nt - 1))
        gx = torch.stack(torch.meshgrid(torch.linspace(0, L1, H), torch.linspace(0, L2, W), indexing="ij"))
        self.x = torch.movedim(gx, 0, -1)     # [H, W, 2]

    def __repr__(self):
        return self.name


class ElementWise(nn.Module):
    """Two 1x1 convs over cat[u (c*tw), pos (nd), conditioning] (enc_grid.py:24-50)."""

    def __init__(self, pde, num_c, num_spatial_dims, time_window, hidden_features, n_cond, activation=None, **kwargs):
        super().__init__()
        act = activation if activation is not None else nn.SiLU()
        cin = num_c * time_window + num_spatial_dims + n_cond
        self.encoder = nn.Sequential(nn.Conv2d(cin, hidden_features, kernel_size=1), act,
                                     nn.Conv2d(hidden_features, hidden_features, kernel_size=1), act)

    def forward(self, u, pos, variables_broadcast=None, **kwargs):
        parts = [u.flatten(1, 2), torch.movedim(pos, -1, 1)]
        if variables_broadcast is not None:
            parts.append(variables_broadcast)
        h = torch.cat(parts, dim=1)
        for layer in self.encoder:                                # 1x1 convs on the tensor-core GEMM kernels
            h = ops.conv1x1(h, layer) if isinstance(layer, nn.Conv2d) else layer(h)
        return h


class TimeConvDense(nn.Module):
    """1x1 conv to 3*tw channels, then a small Conv1d stack over that axis per pixel, then
    u_last + cumsum(dt) * delta (dec_grid.py:97-146, add_delta :8-23)."""

    def __init__(self, pde, num_c, num_spatial_dims, time_window, hidden_features, activation,
                 dec_delta_mode='per_step', dec_delta_dt=True, **kwargs):
        super().__init__()
        if dec_delta_mode != 'per_step':
            raise NotImplementedError("only dec_delta_mode='per_step' (the twophase configs) is implemented")
        self.pde, self.num_c, self.time_window = pde, num_c, time_window
        self.dec_delta_dt = dec_delta_dt
        self.pre_decoder = nn.Conv2d(hidden_features, time_window * 3 * num_c, kernel_size=1)
        ka = math.ceil(time_window / 2)
        kb = math.ceil(time_window / 4) + 1 + (1 if time_window % 4 == 0 else 0)
        self.decoder = nn.Sequential(nn.Conv1d(num_c, num_c * 2, ka, stride=2), activation,
                                     nn.Conv1d(num_c * 2, num_c, kb, stride=1))

    def delta(self, h):
        z = ops.conv1x1(h, self.pre_decoder)                     # [b, 3*tw*c, H, W]
        return ops.timeconv_decoder(z, self.decoder, self.num_c, self.time_window)   # [b, c, tw, H, W]

    def steps(self, device):
        """cumsum(dt) over the time window, exactly as add_delta builds it (dec_grid.py:8-23)."""
        dt = self.pde.dt if self.dec_delta_dt else 1
        return torch.cumsum(torch.full((1, 1, self.time_window, 1, 1), dt, device=device), dim=2)

    def forward(self, h, u, **kwargs):
        return u[:, :, -1:] + self.steps(h.device) * self.delta(h)


_REGISTRY = {"FNO": FNO, "UFNO": UFNO, "UNetModern": UNetModern,
             "enc_grid.ElementWise": ElementWise, "dec_grid.TimeConvDense": TimeConvDense}


def create_model(spec, pde, base_args, extra_kwargs=None):
    """Name lookup like the reference's create_model (enc_proc_dec.py:14-38): an nn.Module is taken as is,
    a string / dict(object=..., **overrides) is resolved against the registry."""
    if isinstance(spec, nn.Module):
        return spec
    if isinstance(spec, str):
        name, kw = spec, dict(base_args)
    elif isinstance(spec, dict):
        spec = dict(spec)
        name = spec.pop("object")
        kw = {**base_args, **spec}
    else:
        raise ValueError("Model was not the correct type: Should be nn.Module / dict / str")
    if extra_kwargs:
        kw.update(extra_kwargs)
    if name not in _REGISTRY:
        raise ValueError(f"Cannot find object {name} in any of {sorted(_REGISTRY)}")
    return _REGISTRY[name](**kw, pde=pde)


class EncProcDec(nn.Module):
    """encoder -> processor(s) -> decoder on a grid (enc_proc_dec.py:41-183, grid branch only)."""

    def __init__(self, pde, encoder, processor, decoder, bc_encoder=None, num_c=1, num_spatial_dims=1, time_window=25,
                 data_structure="grid", processor_residual=False, **base_args):
        super().__init__()
        if data_structure != "grid":
            raise NotImplementedError("only data_structure='grid' is implemented (the GNN path is deprecated upstream)")
        if bc_encoder is not None:
            raise NotImplementedError("bc_encoder is not used by the twophase configs")
        self.pde, self.num_c, self.num_spatial_dims, self.time_window = pde, num_c, num_spatial_dims, time_window
        self.processor_residual = processor_residual
        self.bc_encoder = None
        self.n_cond = pde.n_cond_static + pde.n_cond_spatial
        base_args = dict(base_args, num_c=num_c, num_spatial_dims=num_spatial_dims, time_window=time_window,
                         n_cond=self.n_cond)
        self.encoder = create_model(encoder, pde, base_args)
        procs = processor if isinstance(processor, (list, tuple)) else [processor]
        self.processor = nn.ModuleList([create_model(p, pde, base_args) for p in procs])
        self.decoder = create_model(decoder, pde, base_args)

    @property
    def model_interface(self):
        mi = [p.model_interface for p in self.processor]
        assert mi.count(mi[0]) == len(mi), "Not all processors have the same model interface!"
        return mi[0]

    @property
    def data_interface(self):
        return set.intersection(*[set(p.data_interface) for p in self.processor])

    @staticmethod
    def _none_if_empty(t):
        return None if (t is None or t.numel() == 0) else t

    def conditioning(self, u, cond, spatial_cond):
        """[b, n_cond, H, W]: static scalars broadcast over the grid, then the spatial mask (enc_proc_dec.py:126-137)."""
        cond, spatial_cond = self._none_if_empty(cond), self._none_if_empty(spatial_cond)
        vb = None
        if cond is not None:
            vb = cond[:, :, None, None].expand(-1, -1, *u.shape[3:]).contiguous()
        if spatial_cond is not None:
            vb = spatial_cond if vb is None else torch.cat([vb, spatial_cond], dim=1)
        return vb

    def forward(self, x, cond=None, bc=None, pos=None, t_cond=None, spatial_cond=None):
        if self._none_if_empty(bc) is not None or self._none_if_empty(t_cond) is not None:
            raise NotImplementedError("time-varying conditioning needs a bc_encoder, which the twophase configs do not use")
        vb = self.conditioning(x, cond, spatial_cond)
        h = self.encoder(u=x, variables_broadcast=vb, pos=pos)
        for i, p in enumerate(self.processor):
            nxt = p(h=h, variables_broadcast=vb, pos=pos)
            h = nxt + h if (self.processor_residual and i > 0) else nxt
        return self.decoder(h=h, u=x, variables=None, variables_broadcast=vb, pos=pos)


class ConstrainedSurrogate(EncProcDec):
    """EncProcDec + the output constraints of the reference's `activation_wrapper` (activation_wrapper.py:9-108):
    final activation, obstacle masking, and the 'individual_static' approximate volume preservation."""

    def __init__(self, activation_final, enforce_spatial_cond=False, spatial_cond_channel=0,
                 approx_volume_preserve=False, approx_volume_preserve_mode='block', max_pct_dif=1, **kwargs):
        super().__init__(**kwargs)
        if approx_volume_preserve and approx_volume_preserve_mode != 'individual_static':
            raise NotImplementedError("only approx_volume_preserve_mode='individual_static' (cfg_twophase_*) is implemented")
        self.activation_final = activation_final
        self.enforce_spatial_cond = enforce_spatial_cond
        self.spatial_cond_channel = spatial_cond_channel
        self.approx_volume_preserve = approx_volume_preserve
        self.max_pct_dif = max_pct_dif

    def _mask_out(self, spatial_cond, u):
        m = spatial_cond[:, self.spatial_cond_channel][:, None, None]
        return u - m * u

    def _fused_ok(self, x, spatial_cond):
        return (ops.enable_fused_constraints and not torch.is_grad_enabled() and x.is_cuda and x.dtype == torch.float32
                and self.num_c == 1 and x.dim() == 5 and isinstance(self.decoder, TimeConvDense)
                and isinstance(self.activation_final, (nn.Tanh, nn.Identity))
                and (not self.enforce_spatial_cond or (spatial_cond is not None and spatial_cond.numel() > 0
                                                       and self.spatial_cond_channel == 0)))

    def _consts(self, device, tw):
        key = (str(device), tw)
        if getattr(self, "_const_key", None) != key:
            steps = self.decoder.steps(device).reshape(-1).contiguous()
            cap = torch.cumsum(torch.full((tw,), float(self.max_pct_dif), device=device, dtype=torch.float32), dim=0)   # :86-88
            self._const_key, self._const_val = key, (steps, cap)
        return self._const_val

    def forward(self, x, cond=None, bc=None, pos=None, t_cond=None, spatial_cond=None):
        if self._fused_ok(x, spatial_cond):
            # no-grad application (rollout, push-forward unroll): add_delta + tanh + masking + volume rescale in ONE kernel
            if self._none_if_empty(bc) is not None or self._none_if_empty(t_cond) is not None:
                raise NotImplementedError("time-varying conditioning needs a bc_encoder, which the twophase configs do not use")
            vb = self.conditioning(x, cond, spatial_cond)
            h = self.encoder(u=x, variables_broadcast=vb, pos=pos)
            for i, p in enumerate(self.processor):
                nxt = p(h=h, variables_broadcast=vb, pos=pos)
                h = nxt + h if (self.processor_residual and i > 0) else nxt
            delta = self.decoder.delta(h)
            steps, cap = self._consts(x.device, x.shape[2])
            return ops.constrain_forward(delta, x, spatial_cond if self.enforce_spatial_cond else None, steps, cap,
                                         isinstance(self.activation_final, nn.Tanh), self.enforce_spatial_cond,
                                         self.approx_volume_preserve)
        u = self.activation_final(super().forward(x, cond=cond, bc=bc, pos=pos, t_cond=t_cond, spatial_cond=spatial_cond))
        if self.enforce_spatial_cond:
            u = self._mask_out(spatial_cond, u)
        if self.approx_volume_preserve:
            tw = u.shape[2]
            new_tot = u.sum(dim=(3, 4))                                        # [b, c, tw]
            prev_tot = x[:, :, -1].sum(dim=(2, 3))[:, :, None].expand(-1, -1, tw)
            cap = torch.cumsum(torch.full_like(new_tot, self.max_pct_dif), dim=2)
            dif = (1 - new_tot / prev_tot) * 100
            dif = torch.tanh(dif / cap) / 100 * cap
            u = (u / new_tot[..., None, None]) * ((1 - dif) * prev_tot)[..., None, None]
            if self.enforce_spatial_cond:
                u = self._mask_out(spatial_cond, u)
        return u


def twophase_model_kwargs(processor="UFNO", hidden_features=192, fno_modes=10, hidden_blocks=3):
    """The `model` dict of cfg_twophase_ufno.py:51-89 (processor='UFNO') or cfg_twophase_ufno_fno.py:51-90
    (processor=[dict(object='FNO', hidden_blocks=1), dict(object='UFNO', hidden_blocks=1)])."""
    return dict(
        activation_final=nn.Tanh(), enforce_spatial_cond=True, spatial_cond_channel=0, approx_volume_preserve=True,
        approx_volume_preserve_mode='individual_static', max_pct_dif=1 / 25,
        num_c=1, num_spatial_dims=2, time_window=25, data_structure="grid", processor_residual=False,
        encoder="enc_grid.ElementWise", activation=nn.GELU(), processor=copy.deepcopy(processor), fno_modes=fno_modes,
        hidden_blocks=hidden_blocks, hidden_features=hidden_features, fno_kernel_size=1, fno_conv_mode="single",
        padding_mode="circular", ch_mults=[1, 1], is_attn=[False, False], mid_attn=False, norm=True, use1x1=True,
        decoder="dec_grid.TimeConvDense", dec_delta_mode='per_step')


def build_twophase_model(pde=None, **overrides):
    pde = pde if pde is not None else TwoPhasePDE()
    kw = twophase_model_kwargs()
    kw.update(overrides)
    return ConstrainedSurrogate(**kw, pde=pde)
