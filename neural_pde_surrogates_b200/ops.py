"""torch.autograd binding of the sm_100a spectral-block kernels (the only compute path: no fallback).

`fno_block(h, vb, res, w1, w2, wc, bias, act)` computes, in one chain of hand-written kernels,

    act( SpectralConv2d(cat[h, vb]) + Conv2d_1x1(cat[h, vb]) + bias + res )

which is the body of reference FNO_Layer.forward (proc_fno.py:133-155) plus the U-FNO block tail
`activation(h_fno + h_unet)` (proc_ufno.py:111-118).  PyTorch is used for device memory, streams and autograd
bookkeeping only; every arithmetic pass over the data is a kernel from csrc/.
"""
from __future__ import annotations

import weakref

import numpy as np
import torch

from . import _native

ACT_NONE, ACT_GELU = _native.ACT_NONE, _native.ACT_GELU

_tables_cache: dict = {}


class _PackCache:
    """Tensor-core operands (hi/lo TF32 split, UMMA canonical layout) packed from a weight, reused until the weight
    changes.  Keyed by the Parameter OBJECT through a weak reference (never by address alone: the caching allocator
    hands a freed model's addresses to the next one) and validated by `_version`, `data_ptr`, shape and a global epoch.
    In-place updates (optimizer steps, `load_state_dict`, `copy_`) bump `_version`; writes that bypass the version
    counter (`p.data.copy_()`, `dist.broadcast(p.data)`) must be followed by `invalidate_weight_caches()` -- dp.py
    does so.  During CUDA-graph capture the cache is bypassed (the pack kernel is captured with the graph), so a
    replay always sees the current weights."""

    def __init__(self):
        self._d = {}
        self.epoch = 0

    def get(self, param, kind, build, also=()):
        """`also`: further parameters the packed operand depends on (their versions join the validity stamp)."""
        if param is None or torch.cuda.is_current_stream_capturing():
            return None
        key = id(param)
        stamp = tuple((q._version, q.data_ptr(), tuple(q.shape)) for q in (param, *also)) + (self.epoch,)
        ent = self._d.get(key)
        if ent is None or ent[0]() is not param or ent[1] != stamp:
            ent = (weakref.ref(param, lambda _r, k=key, d=self._d: d.pop(k, None)), stamp, {})
            self._d[key] = ent
        t = ent[2].get(kind)
        if t is None:
            t = build()
            ent[2][kind] = t
        return t


_pack_cache = _PackCache()


def invalidate_weight_caches():
    """Forget every packed weight operand (call after writing parameters through `.data` or a raw pointer)."""
    _pack_cache.epoch += 1
    _pack_cache._d.clear()

# ---- instrumentation used by bench.py: kernel-launch counter and in-situ CUDA-event timing of the two chains
_counters = {"launches": 0}
_timing = {"on": False, "block_forward": [], "block_backward": []}


def reset_counters():
    _counters["launches"] = 0


def counters():
    return dict(_counters)


def add_launches(n: int):
    """Account for kernels launched by a CUDA-graph replay (the Python-side counter does not run during a replay)."""
    _counters["launches"] += int(n)


def enable_timing(flag: bool):
    _timing["on"] = bool(flag)
    if flag:
        _timing["block_forward"], _timing["block_backward"] = [], []


def collect_timings():
    """Milliseconds per recorded chain launch (synchronises)."""
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in _timing[k]] for k in ("block_forward", "block_backward")}


class _Timed:
    def __init__(self, key):
        self.key = key

    def __enter__(self):
        if _timing["on"]:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *a):
        if _timing["on"]:
            self.e1.record()
            _timing[self.key].append((self.e0, self.e1))


def _lib():
    return _native.library()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def tables_for(device: torch.device, H: int, W: int, m1: int, m2: int) -> torch.Tensor:
    """Device copy of the twiddle tables for one (H, W, m1, m2); built once in float64 on the host."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), H, W, m1, m2)
    t = _tables_cache.get(key)
    if t is None:
        lib = _lib()
        n = lib.pdes_tables_floats(H, W, m1, m2)
        if n == 0:
            raise ValueError(f"invalid spectral shape H={H} W={W} modes=({m1},{m2})")
        buf = np.zeros(n, dtype=np.float32)
        _native.check(lib, lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data))
        t = torch.from_numpy(buf).to(device)
        _tables_cache[key] = t
    return t


def _pack_1x1(weight, K, N, transposed: bool, col_off: int = 0):
    """pdes_gemm_tc_pack(_t) of a 1x1-conv weight [Cout][Cin(,1,1)] into a fresh buffer.
    transposed=True : operand At[k = i][n = o] = w[o][i]                       (forward GEMM, K = Cin, N = Cout)
    transposed=False: operand At[k = o][n = i] = w[o][col_off + i], i < N      (input-gradient GEMM, K = Cout)"""
    lib = _lib()
    Cout = weight.shape[0]
    w2 = weight.detach().reshape(Cout, -1)
    if not w2.is_contiguous():
        w2 = w2.contiguous()
    Cin = w2.shape[1]
    pack = torch.empty(lib.pdes_gemm_tc_pack_floats(K, N), dtype=torch.float32, device=weight.device)
    if transposed:
        _native.check(lib, lib.pdes_gemm_tc_pack_t(w2.data_ptr(), Cin, K, N, pack.data_ptr(), _stream()))
    else:
        _native.check(lib, lib.pdes_gemm_tc_pack(w2.data_ptr() + 4 * col_off, Cin, K, N, pack.data_ptr(), _stream()))
    _counters["launches"] += 1
    return pack


def _pack_spectral(w1, w2, H):
    """pdes_mix_tc_pack: the packed master copy Wp[m][o][i_pad] of weights1 / weights2 for K2 on tcgen05."""
    lib = _lib()
    Cin, Cout, m1, m2 = w1.shape
    a, b = w1.detach().contiguous(), w2.detach().contiguous()
    wp = torch.empty(lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2), dtype=torch.float32, device=w1.device)
    _native.check(lib, lib.pdes_mix_tc_pack(a.data_ptr(), b.data_ptr(), wp.data_ptr(), Cin, Cout, H, m1, m2, _stream()))
    _counters["launches"] += 1
    _counters["spectral_packs"] = _counters.get("spectral_packs", 0) + 1
    return wp


# K2 on the tensor cores from the packed master copy (default); False keeps the FFMA kernels on the parameter layout
enable_mix_tc = True


def _tc_pack_usable(weight) -> bool:
    return (weight is not None and weight.is_cuda and weight.dtype == torch.float32
            and _lib().pdes_get_tensor_core_mode() >= 1)


def _check_f32_cuda(name, t, ndim=None):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"{name} is on {t.device}: the B200 spectral block only runs on CUDA (no CPU fallback)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (the reference path is fp32-only, proc_fno.py:265), got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")


class FNOBlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, vb, res, w1, w2, wc, bias, act, wpack_f=None, wpack_b=None, grad_on=True, wspec=None):
        lib = _lib()
        _check_f32_cuda("h", h, 4)
        _check_f32_cuda("variables_broadcast", vb, 4)
        _check_f32_cuda("residual", res, 4)
        _check_f32_cuda("w.weight", wc)
        _check_f32_cuda("w.bias", bias)
        if w1.dtype != torch.complex64 or w2.dtype != torch.complex64:
            raise TypeError("spectral weights must be complex64 (proc_fno.py:240-243)")
        B, C0, H, W = h.shape
        C1 = 0 if vb is None else vb.shape[1]
        Cin, Cout, m1, m2 = w1.shape
        if C0 + C1 != Cin:
            raise ValueError(f"input has {C0}+{C1} channels but the spectral weights expect {Cin}")
        # same checks as the asserts in FNO_Layer.forward (proc_fno.py:135-139)
        assert m2 <= W // 2 + 1, 'modes should be at most the spatial dim // 2 + 1 for the last spatial dimension!'
        assert m1 <= H, 'modes should be at most the spatial dim all but the last spatial dimensions!'
        h = h.contiguous()
        vb = None if vb is None else vb.contiguous()
        res = None if res is None else res.contiguous()
        w1c, w2c = w1.contiguous(), w2.contiguous()
        dev = h.device
        # needs_input_grad ignores grad mode (and grad mode is always off inside forward): the wrapper passes the caller's
        # grad mode, so under torch.no_grad() (rollout, push-forward unroll) nothing is saved and K3b does not store
        # the pre-activation
        needs_grad = bool(grad_on) and any(ctx.needs_input_grad)
        with torch.cuda.device(dev):
            tab = tables_for(dev, H, W, m1, m2)
            wc2 = None
            if wc is not None:
                wc2 = wc.reshape(Cout, Cin)
                if not wc2.is_contiguous():
                    wc2 = wc2.contiguous()
            X = torch.empty(B, Cin, 2 * m1, m2, dtype=torch.complex64, device=dev)
            out = torch.empty(B, Cout, H, W, dtype=torch.float32, device=dev)
            pre = torch.empty_like(out) if (needs_grad and act != ACT_NONE) else None
            ws = torch.empty(lib.pdes_block_fwd_workspace_floats(B, Cin, Cout, H, W, m1, m2), dtype=torch.float32, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            with _Timed("block_forward"):
                _native.check(lib, lib.pdes_block_forward(
                    p(h), C0, p(vb), C1, p(w1c), p(w2c), p(wspec), p(wc2), p(wpack_f), p(bias), p(res), p(tab), p(X), p(ws), p(out),
                    p(pre), B, Cout, H, W, m1, m2, act, _stream()))
            _counters["launches"] += 4 + (1 if (wc is not None and wpack_f is None) else 0)   # K1, K2, K3a, K3b (+ pack)
        if needs_grad:
            ctx.save_for_backward(h, vb, w1c, w2c, wc, X, pre, wpack_b, wspec)    # wspec: dX reads the forward pack too
            ctx.act = act
            ctx.has_res = res is not None
            ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        h, vb, w1, w2, wc, X, pre, wpack_b, wspec = ctx.saved_tensors
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("gradient w.r.t. the broadcast conditioning channels is not implemented "
                                      "(they never require grad in the twophase configs, enc_proc_dec.py:126-137)")
        B, C0, H, W = h.shape
        C1 = 0 if vb is None else vb.shape[1]
        Cin, Cout, m1, m2 = w1.shape
        g = g.contiguous()
        dev = h.device
        with torch.cuda.device(dev):
            tab = tables_for(dev, H, W, m1, m2)
            act = ctx.act
            g_pre = torch.empty_like(g) if act != ACT_NONE else None
            dh = torch.empty_like(h)
            gw1, gw2 = torch.empty_like(w1), torch.empty_like(w2)
            wc2 = dwc = dbias = None
            if wc is not None:
                wc2 = wc.reshape(Cout, Cin)
                if not wc2.is_contiguous():
                    wc2 = wc2.contiguous()
                dwc = torch.empty(Cout, Cin, dtype=torch.float32, device=dev)
                dbias = torch.empty(Cout, dtype=torch.float32, device=dev) if ctx.has_bias else None
            ws = torch.empty(lib.pdes_block_bwd_workspace_floats(B, C0, C1, Cout, H, W, m1, m2), dtype=torch.float32, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            with _Timed("block_backward"):
                _native.check(lib, lib.pdes_block_backward(
                    p(g), p(pre), p(h), C0, p(vb), C1, p(X), p(w1), p(w2), p(wspec), p(wc2), p(wpack_b), p(tab), p(ws), p(g_pre), p(dh),
                    p(gw1), p(gw2), p(dwc), p(dbias), B, Cout, H, W, m1, m2, act, _stream()))
            # act_bwd, K1(g), mix_dw, mix_dx, K3a, K3b, wgrad + its reduce
            _counters["launches"] += 5 + (1 if act != ACT_NONE else 0) + (2 if wc is not None else 0) + \
                (1 if (wc is not None and wpack_b is None) else 0)
        d_res = (g_pre if act != ACT_NONE else g) if ctx.has_res else None
        return dh, None, d_res, gw1, gw2, (None if dwc is None else dwc.view_as(wc)), dbias, None, None, None, None, None


def fno_block(h, vb, res, w1, w2, wc, bias, act: int = ACT_NONE):
    """act(spectral(cat[h,vb]) + conv1x1(cat[h,vb]) + bias + res); any of vb / res / wc / bias may be None."""
    wpack_f = wpack_b = None
    if wc is not None and _tc_pack_usable(wc):
        Cout, Cin, C0 = wc.shape[0], w1.shape[0], h.shape[1]
        wpack_f = _pack_cache.get(wc, ("f", Cin, Cout), lambda: _pack_1x1(wc, Cin, Cout, True))
        if torch.is_grad_enabled() and (h.requires_grad or wc.requires_grad or w1.requires_grad):
            wpack_b = _pack_cache.get(wc, ("b", Cout, C0, 0), lambda: _pack_1x1(wc, Cout, C0, False))
    wspec = None
    if enable_mix_tc and w1.is_cuda and h.is_cuda and h.dim() == 4:
        lib = _lib()
        Cin_, Cout_, m1, m2 = w1.shape
        if lib.pdes_mix_tc_ok(h.shape[0], Cin_, Cout_, m1, m2):
            H = h.shape[2]
            with torch.cuda.device(h.device):
                wspec = _pack_cache.get(w1, ("spec", H), lambda: _pack_spectral(w1, w2, H), also=(w2,))
                if wspec is None:                                       # CUDA-graph capture: the pack is part of the graph
                    wspec = _pack_spectral(w1, w2, H)
    return FNOBlockFunction.apply(h, vb, res, w1, w2, wc, bias, act, wpack_f, wpack_b, torch.is_grad_enabled(), wspec)


def act_code(module) -> int | None:
    """Map an activation module to a kernel epilogue code; None means "apply it with torch after the kernel"."""
    if module is None:
        return ACT_NONE
    if isinstance(module, torch.nn.GELU) and getattr(module, "approximate", "none") == "none":
        return ACT_GELU
    if isinstance(module, torch.nn.Identity):
        return ACT_NONE
    return None


# ---------------------------------------------------------------------------------------------------------------
# U-Net branch helper: fused GroupNorm + activation (csrc/groupnorm.cu)
class GroupNormActFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, groups, eps, act):
        lib = _lib()
        _check_f32_cuda("x", x)
        x = x.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // (B * C)
        dev = x.device
        with torch.cuda.device(dev):
            y = torch.empty_like(x)
            stats = torch.empty(B * groups * 2, dtype=torch.float32, device=dev)
            ws = torch.empty((lib.pdes_gn_workspace_bytes(B, C, HW, groups) + 7) // 8, dtype=torch.float64, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            _native.check(lib, lib.pdes_gn_act_forward(p(x), p(weight), p(bias), float(eps), p(y), p(stats), p(ws),
                                                       B, C, HW, groups, act, _stream()))
            _counters["launches"] += 3
        ctx.save_for_backward(x, weight, bias, stats)
        ctx.groups, ctx.act = groups, act
        return y

    @staticmethod
    def backward(ctx, dy):
        lib = _lib()
        x, weight, bias, stats = ctx.saved_tensors
        dy = dy.contiguous()
        B, C = x.shape[0], x.shape[1]
        HW = x.numel() // (B * C)
        dev = x.device
        with torch.cuda.device(dev):
            dx = torch.empty_like(x)
            dw = torch.empty_like(weight) if weight is not None else None
            db = torch.empty_like(bias) if bias is not None else None
            ws = torch.empty((lib.pdes_gn_workspace_bytes(B, C, HW, ctx.groups) + 7) // 8, dtype=torch.float64, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            _native.check(lib, lib.pdes_gn_act_backward(p(dy), p(x), p(weight), p(bias), p(stats), p(dx), p(dw), p(db),
                                                        p(ws), B, C, HW, ctx.groups, ctx.act, _stream()))
            _counters["launches"] += 3
        return dx, dw, db, None, None, None


def group_norm_act(x, norm, act):
    """act(norm(x)) for an nn.GroupNorm `norm` (or Identity) -- one fused kernel pair on CUDA, plain PyTorch otherwise
    (the U-Net branch is a PyTorch component; only its GroupNorm+GELU is replaced on the GPU)."""
    code = act_code(act)
    if isinstance(norm, torch.nn.GroupNorm) and x.is_cuda and x.dtype == torch.float32 and code is not None:
        return GroupNormActFunction.apply(x, norm.weight, norm.bias, norm.num_groups, norm.eps, code)
    y = norm(x)
    return y if act is None else act(y)


# ---------------------------------------------------------------------------------------------------------------
# U-Net branch helper: 3x3 valid convolution, forward on tcgen05 (csrc/gemm_tc.cu: k_conv3x3_tc), backward on cuDNN
class Conv3x3ValidFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib()
        x = x.contiguous()
        w = weight.contiguous()
        B, Cin, H, W = x.shape
        N = w.shape[0]
        dev = x.device
        with torch.cuda.device(dev):
            out = torch.empty(B, N, H - 2, W - 2, dtype=torch.float32, device=dev)
            wpack = torch.empty(lib.pdes_conv3x3_tc_pack_floats(Cin, N), dtype=torch.float32, device=dev)
            p = lambda t: None if t is None else t.data_ptr()
            _native.check(lib, lib.pdes_conv3x3_tc(p(x), p(w), p(bias), p(wpack), p(out), B, Cin, N, H, W, _stream()))
            _counters["launches"] += 2
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        bias_sizes = [w.shape[0]] if ctx.has_bias else None
        mask = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_bias and ctx.needs_input_grad[2]]
        gx, gw, gb = torch.ops.aten.convolution_backward(g.contiguous(), x, w, bias_sizes, [1, 1], [0, 0], [1, 1], False,
                                                         [0, 0], 1, mask)
        return gx, gw, gb


class ConvValidDgradAsForwardFunction(torch.autograd.Function):
    """Stride-1 valid convolution on cuDNN whose input gradient is computed as a FORWARD convolution of the zero-padded
    output gradient with the flipped, transposed filter.  Same arithmetic as cuDNN's dgrad (fp32, TF32 off), but it lands
    on cuDNN's forward engines (Winograd, ~150 TFLOP/s fp32-equivalent at the U-Net shapes on B200) instead of its dgrad
    engines (~50 TFLOP/s): measured 2.7 ms -> 0.9 ms for the 385->192 conv at 100x68, batch 16 (tools/time_conv_bwd.py).
    Weight and bias gradients stay on cuDNN's wgrad (~120 TFLOP/s)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return torch.nn.functional.conv2d(x, weight, bias)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g = g.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            kh, kw = w.shape[-2:]
            wt = w.detach().flip(2, 3).transpose(0, 1).contiguous()
            gx = torch.nn.functional.conv2d(g, wt, None, padding=(kh - 1, kw - 1))
        need_b = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[1] or need_b:
            _, gw, gb = torch.ops.aten.convolution_backward(g, x, w, [w.shape[0]] if ctx.has_bias else None, [1, 1], [0, 0], [1, 1],
                                                            False, [0, 0], 1, [False, True, bool(need_b)])
        return gx, gw, gb


# cuDNN's fp32 dgrad engines are ~3x slower than its forward engines at the U-Net shapes: route dx through a forward conv
enable_dgrad_as_forward = True


def conv3x3_valid(x, conv):
    """conv(x) for the U-Net's 3x3 valid convolutions: tcgen05 implicit GEMM (3xTF32) when the shape allows, else cuDNN."""
    if (x.is_cuda and x.dtype == torch.float32 and isinstance(conv, torch.nn.Conv2d) and tuple(conv.kernel_size) == (3, 3)
            and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) == (1, 1) and conv.groups == 1
            and (tuple(conv.padding) == (0, 0) if not isinstance(conv.padding, str) else conv.padding == "valid")
            and x.dim() == 4 and enable_conv_tc):
        lib = _lib()
        B, Cin, H, W = x.shape
        xc = x if x.is_contiguous() else x.contiguous()
        if lib.pdes_get_tensor_core_mode() >= 2 and lib.pdes_conv3x3_tc_ok(B, Cin, conv.out_channels, H, W, xc.data_ptr()):
            return Conv3x3ValidFunction.apply(xc, conv.weight, conv.bias)
    if (enable_dgrad_as_forward and x.is_cuda and x.dim() == 4 and isinstance(conv, torch.nn.Conv2d)
            and tuple(conv.stride) == (1, 1) and tuple(conv.dilation) == (1, 1) and conv.groups == 1
            and not isinstance(conv.padding, str) and tuple(conv.padding) == (0, 0) and torch.is_grad_enabled()
            and (x.requires_grad or conv.weight.requires_grad)):
        return ConvValidDgradAsForwardFunction.apply(x, conv.weight, conv.bias)
    return conv(x)


# Off by default: measured on B200 (tools/time_conv.py) cuDNN's fp32 Winograd is within 1.2-1.5x of this kernel and more
# accurate (the fp32 accumulation of tcgen05 truncates, so 3xTF32 reaches 1.4e-5 .. 2.6e-5 rel. L2 at K = Cin*9 > 1700,
# above the 1e-5 bar); kept as an opt-in experiment for the next round (split-K over several TMEM accumulators).
enable_conv_tc = False


# ---- U-Net branch: 1x1 convolutions on the K3b / weight-gradient tensor-core kernels ---------
enable_conv1x1_tc = True
_C1_MAX_N = 208          # output channels per launch (tensor-memory A-operand path of the K3b kernel)
_C1_MAX_WG = 255         # input channels per weight-gradient launch (K + 1 ones column <= 256)


def _ranges(total: int, cap: int):
    parts = -(-total // cap)
    size = -(-total // parts)
    return [(o, min(size, total - o)) for o in range(0, total, size)]


class Conv1x1Function(torch.autograd.Function):
    """y = conv2d(x, w[N,Cin,1,1], b) and its three gradients on tcgen05 (3xTF32), replacing cuDNN's SIMT sgemm for the
    shortcut convs of the reference's ResidualBlock (proc_unet_modern.py:219-222) and the encoder/decoder 1x1 convs."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        lib = _lib()
        B, Cin, H, W = x.shape
        N, HW = weight.shape[0], H * W
        dev = x.device
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            pack = _pack_cache.get(weight, ("f", Cin, N), lambda: _pack_1x1(weight, Cin, N, True))
            if pack is None:                                                        # graph capture: pack in-graph
                pack = _pack_1x1(weight, Cin, N, True)
            out = torch.empty(B, N, H, W, dtype=torch.float32, device=dev)
            st = _stream()
            _native.check(lib, lib.pdes_conv1x1_tc(p(x), Cin, p(pack), p(bias), None, p(out), N * HW, B, N, HW, ACT_NONE, st))
            _counters["launches"] += 1
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, g):
        lib = _lib()
        x, weight = ctx.saved_tensors
        B, Cin, H, W = x.shape
        N, HW = weight.shape[0], H * W
        dev = x.device
        g = g.contiguous()
        p = lambda t: None if t is None else t.data_ptr()
        gx = gw = gb = None
        with torch.cuda.device(dev):
            st = _stream()
            w2 = weight.detach().reshape(N, Cin)
            if not w2.is_contiguous():
                w2 = w2.contiguous()
            if ctx.needs_input_grad[0]:
                gx = torch.empty_like(x)
                for off, n in _ranges(Cin, _C1_MAX_N):            # dx[:, off:off+n] = w[:, off:off+n]^T g
                    pack = _pack_cache.get(weight, ("b", N, n, off), lambda: _pack_1x1(weight, N, n, False, off))
                    if pack is None:
                        pack = _pack_1x1(weight, N, n, False, off)
                    _native.check(lib, lib.pdes_conv1x1_tc(p(g), N, p(pack), None, None, p(gx) + 4 * off * HW, Cin * HW, B, n, HW,
                                                           ACT_NONE, st))
                    _counters["launches"] += 1
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                gw = torch.empty(N, Cin, dtype=torch.float32, device=dev)
                gb = torch.empty(N, dtype=torch.float32, device=dev) if ctx.has_bias else None
                ws = torch.empty(lib.pdes_wgrad_tc_workspace_floats(N, min(Cin, _C1_MAX_WG)), dtype=torch.float32, device=dev)
                done = False
                for cap in (_C1_MAX_WG, 128, 64):                 # the kernel's shared-memory need grows with N * channels
                    rcs = []
                    for k, (off, n) in enumerate(_ranges(Cin, cap)):
                        rcs.append(lib.pdes_wgrad_tc_range(p(g), p(x), Cin, off, n, p(gw) + 4 * off, Cin, p(gb) if k == 0 else None,
                                                           p(ws), B, N, HW, st))
                        if rcs[-1] != _native.PDES_OK:
                            break
                        _counters["launches"] += 2
                    if rcs[-1] == _native.PDES_OK:
                        done = True
                        break
                    if rcs[-1] != _native.PDES_ERR_UNSUPPORTED:
                        _native.check(lib, rcs[-1])
                if done:
                    gw = gw.view_as(weight)
                else:                                              # shape outside the tensor-core kernel: cuDNN
                    _, gw, gb = torch.ops.aten.convolution_backward(g, x, weight, [N] if ctx.has_bias else None, [1, 1], [0, 0],
                                                                    [1, 1], False, [0, 0], 1, [False, True, ctx.has_bias])
        return gx, gw, gb


def conv1x1(x, conv):
    """conv(x) for a kernel_size=1 nn.Conv2d: tcgen05 GEMM kernels when the shape allows, else the module itself (cuDNN)."""
    if (enable_conv1x1_tc and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and isinstance(conv, torch.nn.Conv2d)
            and tuple(conv.kernel_size) == (1, 1) and tuple(conv.stride) == (1, 1) and conv.groups == 1
            and tuple(conv.dilation) == (1, 1) and not isinstance(conv.padding, str) and tuple(conv.padding) == (0, 0)):
        lib = _lib()
        B, Cin, H, W = x.shape
        N, HW = conv.out_channels, H * W
        xc = x if x.is_contiguous() else x.contiguous()
        if (HW % 16 == 0 and 8 <= N <= 256 and Cin >= 8
                and (conv.bias is None or conv.bias.data_ptr() % 16 == 0)
                and lib.pdes_conv1x1_tc_ok(B, Cin, N, HW, xc.data_ptr())):
            return Conv1x1Function.apply(xc, conv.weight, conv.bias)
    return conv(x)


# ---- decoder: fused per-pixel temporal Conv1d stack of TimeConvDense -----------------------
enable_timeconv = True


class TimeConvFunction(torch.autograd.Function):
    """delta[b, 0, t, h, w] = Conv1d(2->1,k=8)(act(Conv1d(1->2,k=13,s=2)(z[b, :, h, w])))  (reference dec_grid.py:97-146),
    one thread per pixel, no permute copies; the backward recomputes the hidden layer."""

    @staticmethod
    def forward(ctx, z, w1, b1, w2, b2, act):
        lib = _lib()
        B, C3, H, W = z.shape
        tw = C3 // 3
        dev = z.device
        p = lambda t: t.data_ptr()
        ws_ = [t.detach().contiguous() for t in (w1, b1, w2, b2)]
        with torch.cuda.device(dev):
            out = torch.empty(B, 1, tw, H, W, dtype=torch.float32, device=dev)
            _native.check(lib, lib.pdes_timeconv_forward(p(z), p(ws_[0]), p(ws_[1]), p(ws_[2]), p(ws_[3]), p(out), B, H * W, tw,
                                                         act, _stream()))
            _counters["launches"] += 1
        ctx.save_for_backward(z, *ws_)
        ctx.act = act
        return out

    @staticmethod
    def backward(ctx, gy):
        lib = _lib()
        z, w1, b1, w2, b2 = ctx.saved_tensors
        B, C3, H, W = z.shape
        tw = C3 // 3
        dev = z.device
        gy = gy.contiguous()
        p = lambda t: t.data_ptr()
        with torch.cuda.device(dev):
            dz = torch.empty_like(z)
            dw1, db1, dw2, db2 = (torch.empty_like(t) for t in (w1, b1, w2, b2))
            ws = torch.empty(lib.pdes_timeconv_bwd_workspace_floats(B, H * W, tw), dtype=torch.float32, device=dev)
            _native.check(lib, lib.pdes_timeconv_backward(p(z), p(gy), p(w1), p(b1), p(w2), p(b2), p(dz), p(dw1), p(db1), p(dw2),
                                                          p(db2), p(ws), B, H * W, tw, ctx.act, _stream()))
            _counters["launches"] += 2
        return dz, dw1, db1, dw2, db2, None


def timeconv_decoder(z, decoder, num_c: int, time_window: int):
    """decoder = nn.Sequential(Conv1d, act, Conv1d) applied per pixel along the channel axis of z [B, num_c*3*tw, H, W];
    returns [B, num_c, tw, H, W].  Fused kernel for the shipped decoder (tw = 25, one field, GELU / identity), else torch."""
    B, _, H, W = z.shape
    if enable_timeconv and z.is_cuda and z.dtype == torch.float32 and len(decoder) == 3:
        c1, a, c2 = decoder[0], decoder[1], decoder[2]
        code = act_code(a)
        if (code is not None and isinstance(c1, torch.nn.Conv1d) and isinstance(c2, torch.nn.Conv1d)
                and c1.bias is not None and c2.bias is not None and tuple(c1.stride) == (2,) and tuple(c2.stride) == (1,)
                and tuple(c1.padding) == (0,) and tuple(c2.padding) == (0,) and c1.groups == 1 and c2.groups == 1
                and _lib().pdes_timeconv_ok(time_window, num_c)):
            zc = z if z.is_contiguous() else z.contiguous()
            return TimeConvFunction.apply(zc, c1.weight, c1.bias, c2.weight, c2.bias, code)
    zz = z.permute(0, 2, 3, 1).reshape(B * H * W, num_c, time_window * 3)
    return decoder(zz).view(B, H, W, num_c, time_window).permute(0, 3, 4, 1, 2)


# ---- per-step wrapper: fused output constraints (csrc/constrain.cu), inference / no-grad applications only -----------
enable_fused_constraints = True


def constrain_forward(delta, x, mask, steps, cap, use_tanh: bool, use_mask: bool, use_volume: bool):
    """mask(volume_rescale(mask(tanh(x[:, :, -1:] + steps * delta)))) for one field; delta, x: [B, 1, tw, H, W]."""
    lib = _lib()
    _check_f32_cuda("delta", delta, 5)
    _check_f32_cuda("x", x, 5)
    B, C, tw, H, W = delta.shape
    if C != 1 or tuple(x.shape) != tuple(delta.shape):
        raise ValueError("constrain_forward handles one field with matching state / delta shapes")
    delta, x = delta.contiguous(), x.contiguous()
    p = lambda t: None if t is None else t.data_ptr()
    with torch.cuda.device(delta.device):
        out = torch.empty_like(delta)
        mk = None
        if use_mask:
            mk = mask if mask.is_contiguous() else mask.contiguous()
        _native.check(lib, lib.pdes_constrain_forward(p(delta), p(x), p(mk), (mk.shape[1] * H * W) if mk is not None else 0,
                                                      p(steps), p(cap), p(out), B, tw, H * W, int(use_tanh), int(use_mask),
                                                      int(use_volume), _stream()))
        _counters["launches"] += 1
    return out
