"""Parity checks of every C-ABI entry point against the float64 oracle, written once and run on both the CPU
emulation build (tests/test_kernels_emu.py) and the real sm_100a library (tests/test_kernels_gpu.py).

Tolerance: float32 kernels vs float64 oracle, relative L2 <= 1e-5 (BASELINE.json north_star: "fp32 forward
outputs and gradients within relative L2 1e-5 per layer")."""
from __future__ import annotations

import numpy as np

from oracle import spectral_oracle as so

TOL = 1e-5

# (B, C0, C1, Cout, H, W, m1, m2) -- small shapes with the reference's edge cases:
SMALL_SHAPES = [
    (2, 5, 0, 4, 12, 8, 3, 5),     # m2 = W/2+1 (Nyquist column kept), no conditioning channels
    (2, 3, 1, 4, 8, 8, 5, 3),      # 2*m1 > H: second block overwrites first-block rows (proc_fno.py:266-269)
    (1, 6, 2, 6, 16, 16, 4, 4),    # vector paths (W%4==0, M%... not multiple of 4)
    (2, 4, 1, 8, 9, 7, 2, 3),      # odd H, odd W: scalar paths, ragged tiles
    (1, 3, 0, 2, 6, 10, 6, 2),     # m1 = H
    (3, 7, 1, 12, 20, 12, 4, 5),   # HW=240: pixel tile crosses image end, Cout multiple of 4
    (1, 3, 1, 4, 96, 64, 10, 10),  # the shipped grid / modes: takes the specialised K1 fast path
]


def _rng(seed):
    return np.random.default_rng(seed)


def _weights(rng, Cin, Cout, m1, m2):
    def one():
        return (rng.standard_normal((Cin, Cout, m1, m2)) + 1j * rng.standard_normal((Cin, Cout, m1, m2))) / np.sqrt(Cin)
    return one().astype(np.complex64), one().astype(np.complex64)


def check_dft_fwd(be, shape, herm=0):
    B, C0, C1, _, H, W, m1, m2 = shape
    rng = _rng(1)
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    tab = be.tables(H, W, m1, m2)
    dx0, dx1 = be.upload(x0), (be.upload(x1) if C1 else None)
    X = be.empty((B, C0 + C1, 2 * m1, m2), complex_=True)
    be.check(be.lib.pdes_dft_fwd(be.ptr(dx0), C0, be.ptr(dx1), C1, B, H, W, m1, m2, be.ptr(tab), herm, be.ptr(X), be.stream))
    xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    ref = so.dft_fwd_pruned(xin, m1, m2, lscale=so.hermitian_scale(H, W, m2) if herm else None)
    err = so.rel_l2(be.download(X), ref)
    assert err < TOL, f"dft_fwd {shape} herm={herm}: rel L2 {err:.3e}"
    return err


def check_mix(be, shape):
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(2)
    X = (rng.standard_normal((B, Cin, 2 * m1, m2)) + 1j * rng.standard_normal((B, Cin, 2 * m1, m2))).astype(np.complex64)
    GO = (rng.standard_normal((B, Cout, 2 * m1, m2)) + 1j * rng.standard_normal((B, Cout, 2 * m1, m2))).astype(np.complex64)
    w1, w2 = _weights(rng, Cin, Cout, m1, m2)
    dX, dGO, dw1, dw2 = be.upload(X), be.upload(GO), be.upload(w1), be.upload(w2)
    errs = {}
    for nsplit in (1, min(3, Cin)):
        P = be.empty((nsplit, B, Cout, 2 * m1, m2), complex_=True)
        be.check(be.lib.pdes_mix_fwd(be.ptr(dX), be.ptr(dw1), be.ptr(dw2), be.ptr(P), nsplit, B, Cin, Cout, H, m1, m2, be.stream))
        errs[f"fwd{nsplit}"] = so.rel_l2(be.download(P).sum(axis=0), so.mode_mix(X, w1, w2, H))
    for nsplit in (1, min(2, Cout)):
        P = be.empty((nsplit, B, C0, 2 * m1, m2), complex_=True)
        be.check(be.lib.pdes_mix_dx(be.ptr(dGO), be.ptr(dw1), be.ptr(dw2), be.ptr(P), nsplit, B, Cin, Cout, C0, H, m1, m2, be.stream))
        errs[f"dx{nsplit}"] = so.rel_l2(be.download(P).sum(axis=0), so.mode_mix_dx(GO, w1, w2, H)[:, :C0])
    g1 = be.empty((Cin, Cout, m1, m2), complex_=True)
    g2 = be.empty((Cin, Cout, m1, m2), complex_=True)
    be.check(be.lib.pdes_mix_dw(be.ptr(dX), be.ptr(dGO), be.ptr(g1), be.ptr(g2), B, Cin, Cout, H, m1, m2, be.stream))
    r1, r2 = so.mode_mix_dw(X, GO, H, m1)
    errs["dw1"] = so.rel_l2(be.download(g1), r1)
    errs["dw2"] = so.rel_l2(be.download(g2), r2)
    for k, v in errs.items():
        assert v < TOL, f"mix {k} {shape}: rel L2 {v:.3e}"
    return errs


def mt_geom(B, Cin, Cout, m1, m2, sms=148):
    """Python mirror of mt_geom() in csrc/spectral_mix_tc.cu (how K2 on tcgen05 cuts its chunk stream)."""
    CinP = (Cin + 15) // 16 * 16
    ntile = (Cout + 127) // 128
    to = ((Cout + ntile - 1) // ntile + 7) // 8 * 8
    nck = CinP // 16
    nitems = 2 * m1 * m2 * ntile
    nch = nitems * nck
    G = max(min(sms, nitems), 1)
    per = max((nch + G - 1) // G, nck)
    split = np.array([(it * nck) // per != ((it + 1) * nck - 1) // per for it in range(nitems)])
    return dict(CinP=CinP, ntile=ntile, to=to, nck=nck, nitems=nitems, per=per, split=split)


def check_spectral_tc(be, shape, run_mma):
    """The tensor-core K2 family: packed master copy of the weights, K1's mode-major spectrum, K3a on K2's output layout
    (both builds) and, with run_mma (sm_100a only), the tcgen05 mix itself -- all against the float64 oracle."""
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin, MM2 = C0 + C1, 2 * m1 * m2
    lib = be.lib
    rng = _rng(7)
    g = mt_geom(B, Cin, Cout, m1, m2)
    CinP = g["CinP"]
    w1, w2 = _weights(rng, Cin, Cout, m1, m2)
    dw1, dw2 = be.upload(w1), be.upload(w2)
    # ---- packed master copy: Wp[m][o][i_pad] complex, dead rows and pad columns zero
    ntile, to, nck = g["ntile"], g["to"], g["nck"]
    assert lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2) == MM2 * ntile * to * CinP * 2
    Wp = be.empty((MM2, ntile, nck, to, 16), complex_=True)               # [m][tile][chunk][row][16 i]
    be.check(lib.pdes_mix_tc_pack(be.ptr(dw1), be.ptr(dw2), be.ptr(Wp), Cin, Cout, H, m1, m2, be.stream))
    wt = np.concatenate([w1, w2], axis=2) * so.live_rows(H, m1)[None, None, :, None]        # [Cin, Cout, 2m1, m2]
    full = np.zeros((MM2, ntile * to, CinP), dtype=np.complex64)          # [m][o padded][i padded]
    full[:, :Cout, :Cin] = wt.reshape(Cin, Cout, MM2).transpose(2, 1, 0)
    ref_wp = full.reshape(MM2, ntile, to, nck, 16).transpose(0, 1, 3, 2, 4)
    assert np.array_equal(be.download(Wp), ref_wp), f"weight pack {shape}"
    # ---- K1 with the mode-major copy
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    tab = be.tables(H, W, m1, m2)
    dx0, dx1 = be.upload(x0), (be.upload(x1) if C1 else None)
    X = be.empty((B, Cin, 2 * m1, m2), complex_=True)
    assert lib.pdes_mix_tc_x2_floats(B, Cin, m1, m2) == MM2 * B * CinP * 2
    X2 = be.empty((MM2, B, CinP), complex_=True)
    be.check(lib.pdes_dft_fwd2(be.ptr(dx0), C0, be.ptr(dx1), C1, B, H, W, m1, m2, be.ptr(tab), 0, be.ptr(X), be.ptr(X2), be.stream))
    Xh, X2h = be.download(X), be.download(X2)
    xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    assert so.rel_l2(Xh, so.dft_fwd_pruned(xin, m1, m2)) < TOL
    assert np.array_equal(X2h[:, :, :Cin], Xh.reshape(B, Cin, MM2).transpose(2, 0, 1)), f"mode-major spectrum {shape}"
    # ---- K2 on tcgen05 (or, without a GPU, a synthetic K2 output with the same partial-sum convention)
    O_ref = so.mode_mix(Xh.astype(np.complex128), w1, w2, H)                                 # [B, Cout, 2m1, m2]
    assert lib.pdes_mix_tc_o2_floats(B, Cout, m1, m2) == 2 * MM2 * B * Cout * 2
    item_of = (np.arange(MM2)[:, None] * g["ntile"] + (np.arange(Cout)[None, :] // g["to"]))  # [m][o] -> work item
    split_mo = g["split"][item_of]                                                            # [m][o]
    if run_mma:
        assert lib.pdes_mix_tc_ok(B, Cin, Cout, m1, m2) == 1
        for rep in range(4):                                                                  # repeated: the kernel's roles race if a protocol is wrong
            O2 = be.empty((2, MM2, B, Cout), complex_=True)                                   # NaN prefilled
            be.check(lib.pdes_mix_tc_fwd(be.ptr(X2), be.ptr(Wp), be.ptr(O2), B, Cin, Cout, m1, m2, be.stream))
            O2h = be.download(O2)
            assert (O2h[1][~np.broadcast_to(split_mo[:, None, :], O2h[1].shape)] == 0).all(), "partial 1 of an unsplit item must be zero"
            got = (O2h[0] + O2h[1]).transpose(1, 2, 0).reshape(B, Cout, 2 * m1, m2)
            err = so.rel_l2(got, O_ref)
            assert err < TOL, f"mix_tc {shape} (repeat {rep}): rel L2 {err:.3e}"
    else:
        Om = O_ref.reshape(B, Cout, MM2).transpose(2, 0, 1).astype(np.complex64)              # [m][b][o]
        U = (rng.standard_normal(Om.shape) + 1j * rng.standard_normal(Om.shape)).astype(np.complex64)
        sp = np.broadcast_to(split_mo[:, None, :], Om.shape)
        O2h = np.stack([np.where(sp, Om - U, Om), np.where(sp, U, 0)]).astype(np.complex64)
        O2 = be.upload(O2h)
    # ---- K3a on that layout
    Z = be.empty((B, H, 2 * m2, Cout))
    be.check(lib.pdes_inv_h_modes(be.ptr(O2), B, Cout, H, m1, m2, be.ptr(tab), be.ptr(Z), be.stream))
    kx = so.kx_table(H, m1)
    E = np.exp(2j * np.pi * np.outer(np.arange(H), kx) / H)                                   # [H, 2m1]
    Zc = np.einsum("hk,bokl->bhlo", E, O_ref)                                                 # [B, H, m2, Cout]
    Zref = np.stack([Zc.real, Zc.imag], axis=3).reshape(B, H, 2 * m2, Cout)
    err = so.rel_l2(be.download(Z), Zref)
    assert err < TOL, f"inv_h_modes {shape}: rel L2 {err:.3e}"
    return err


def check_inverse(be, shape, with_gemm=True, act=1, backward_scale=0):
    """K3a + K3b against the oracle: spectral inverse (+ 1x1 conv + bias + residual + GELU)."""
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(3)
    nsplit = 2
    P = (rng.standard_normal((nsplit, B, Cout, 2 * m1, m2)) + 1j * rng.standard_normal((nsplit, B, Cout, 2 * m1, m2))).astype(np.complex64)
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    wc = (rng.standard_normal((Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32)
    res = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    tab = be.tables(H, W, m1, m2)
    dP = be.upload(P)
    Z = be.empty((B, H, 2 * m2, Cout))
    be.check(be.lib.pdes_inv_h(be.ptr(dP), nsplit, B, Cout, H, m1, m2, be.ptr(tab), be.ptr(Z), be.stream))
    dwc = be.upload(wc)
    wct = be.empty((Cin, Cout))
    be.check(be.lib.pdes_transpose(be.ptr(dwc), be.ptr(wct), Cout, Cin, be.stream))
    assert np.array_equal(be.download(wct), wc.T)
    dx0, dx1 = be.upload(x0), (be.upload(x1) if C1 else None)
    dbias, dres = be.upload(bias), be.upload(res)
    out = be.empty((B, Cout, H, W))
    pre = be.empty((B, Cout, H, W))
    if with_gemm:
        be.check(be.lib.pdes_inv_w_gemm(be.ptr(Z), be.ptr(wct), Cout, be.ptr(dx0), C0, be.ptr(dx1), C1, be.ptr(dbias),
                                        be.ptr(dres), be.ptr(tab), backward_scale, be.ptr(out), be.ptr(pre),
                                        B, Cout, H, W, m1, m2, act, be.stream))
    else:
        be.check(be.lib.pdes_inv_w_gemm(be.ptr(Z), None, 0, None, 0, None, 0, None, None, be.ptr(tab), backward_scale,
                                        be.ptr(out), None, B, Cout, H, W, m1, m2, act, be.stream))
    O = P.astype(np.complex128).sum(axis=0)
    lscale = np.ones(m2) if backward_scale else so.hermitian_scale(H, W, m2)
    ref = so.inv_pruned(O, H, W, lscale)
    if with_gemm:
        xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
        ref = ref + np.einsum("oi,bihw->bohw", wc.astype(np.float64), xin) + bias[None, :, None, None] + res
        e_pre = so.rel_l2(be.download(pre), ref)
        assert e_pre < TOL, f"inverse pre {shape}: rel L2 {e_pre:.3e}"
    refo = so.gelu(ref) if act == 1 else ref
    err = so.rel_l2(be.download(out), refo)
    assert err < TOL, f"inverse out {shape} gemm={with_gemm} act={act}: rel L2 {err:.3e}"
    return err


def check_pointwise(be, shape):
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(4)
    g = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    pre = (2.0 * rng.standard_normal((B, Cout, H, W))).astype(np.float32)
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    dg, dpre = be.upload(g), be.upload(pre)
    gp = be.empty((B, Cout, H, W))
    be.check(be.lib.pdes_act_bwd(be.ptr(dg), be.ptr(dpre), be.ptr(gp), g.size, 1, be.stream))
    e1 = so.rel_l2(be.download(gp), g * so.gelu_grad(pre))
    assert e1 < TOL, f"act_bwd {shape}: {e1:.3e}"
    ws = be.empty((be.lib.pdes_wgrad_workspace_floats(B, Cout, Cin, H * W),))
    dW = be.empty((Cout, Cin))
    db = be.empty((Cout,))
    dx0, dx1 = be.upload(x0), (be.upload(x1) if C1 else None)
    be.check(be.lib.pdes_wgrad(be.ptr(dg), be.ptr(dx0), C0, be.ptr(dx1), C1, be.ptr(dW), be.ptr(db), be.ptr(ws),
                               B, Cout, H * W, be.stream))
    xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    e2 = so.rel_l2(be.download(dW), np.einsum("bohw,bihw->oi", g.astype(np.float64), xin))
    e3 = so.rel_l2(be.download(db), g.astype(np.float64).sum(axis=(0, 2, 3)))
    assert e2 < TOL and e3 < TOL, f"wgrad {shape}: dW {e2:.3e} dbias {e3:.3e}"
    return e1, e2, e3


def block_inputs(shape, seed=5, reference_init=False):
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(seed)
    d = dict(
        h=rng.standard_normal((B, C0, H, W)).astype(np.float32),
        vb=(rng.random((B, C1, H, W)) < 0.3).astype(np.float32) if C1 else None,
        res=rng.standard_normal((B, Cout, H, W)).astype(np.float32),
        g=rng.standard_normal((B, Cout, H, W)).astype(np.float32),
        wc=((rng.random((Cout, Cin)) * 2 - 1) / np.sqrt(Cin)).astype(np.float32),
        bias=((rng.random(Cout) * 2 - 1) / np.sqrt(Cin)).astype(np.float32),
    )
    if reference_init:
        # the reference's init: scale * U[0,1) for re and im, scale = 1/(Cin*Cout)  (proc_fno.py:239-243)
        sc = 1.0 / (Cin * Cout)
        d["w1"] = (sc * (rng.random((Cin, Cout, m1, m2)) + 1j * rng.random((Cin, Cout, m1, m2)))).astype(np.complex64)
        d["w2"] = (sc * (rng.random((Cin, Cout, m1, m2)) + 1j * rng.random((Cin, Cout, m1, m2)))).astype(np.complex64)
    else:
        d["w1"], d["w2"] = _weights(rng, Cin, Cout, m1, m2)
    return d


def run_block(be, shape, d, act=1, use_res=True, use_conv=True, spec_pack=False):
    """Forward + backward of the fused chain; returns dict of numpy results.  spec_pack=True hands the chain the
    packed master copy of the spectral weights (=> K2 on tcgen05 where supported)."""
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    lib = be.lib
    tab = be.tables(H, W, m1, m2)
    up = {k: (be.upload(v) if v is not None else None) for k, v in d.items()}
    X = be.empty((B, Cin, 2 * m1, m2), complex_=True)
    ws = be.empty((lib.pdes_block_fwd_workspace_floats(B, Cin, Cout, H, W, m1, m2),))
    wspec = None
    if spec_pack:
        wspec = be.empty((lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2),))
        be.check(lib.pdes_mix_tc_pack(be.ptr(up["w1"]), be.ptr(up["w2"]), be.ptr(wspec), Cin, Cout, H, m1, m2, be.stream))
    out = be.empty((B, Cout, H, W))
    pre = be.empty((B, Cout, H, W))
    be.check(lib.pdes_block_forward(be.ptr(up["h"]), C0, be.ptr(up["vb"]), C1, be.ptr(up["w1"]), be.ptr(up["w2"]),
                                    be.ptr(wspec), be.ptr(up["wc"]) if use_conv else None, None, be.ptr(up["bias"]) if use_conv else None,
                                    be.ptr(up["res"]) if use_res else None, be.ptr(tab), be.ptr(X), be.ptr(ws),
                                    be.ptr(out), be.ptr(pre), B, Cout, H, W, m1, m2, act, be.stream))
    wsb = be.empty((lib.pdes_block_bwd_workspace_floats(B, C0, C1, Cout, H, W, m1, m2),))
    g_pre = be.empty((B, Cout, H, W))
    dh = be.empty((B, C0, H, W))
    gw1 = be.empty((Cin, Cout, m1, m2), complex_=True)
    gw2 = be.empty((Cin, Cout, m1, m2), complex_=True)
    dwc = be.empty((Cout, Cin))
    dbias = be.empty((Cout,))
    be.check(lib.pdes_block_backward(be.ptr(up["g"]), be.ptr(pre), be.ptr(up["h"]), C0, be.ptr(up["vb"]), C1,
                                     be.ptr(X), be.ptr(up["w1"]), be.ptr(up["w2"]), be.ptr(wspec),
                                     be.ptr(up["wc"]) if use_conv else None, None, be.ptr(tab), be.ptr(wsb),
                                     be.ptr(g_pre), be.ptr(dh), be.ptr(gw1), be.ptr(gw2),
                                     be.ptr(dwc) if use_conv else None, be.ptr(dbias) if use_conv else None,
                                     B, Cout, H, W, m1, m2, act, be.stream))
    r = dict(out=be.download(out), pre=be.download(pre), X=be.download(X), dh=be.download(dh),
             dw1=be.download(gw1), dw2=be.download(gw2))
    if act:
        r["dres"] = be.download(g_pre)
    if use_conv:
        r["dwc"] = be.download(dwc)
        r["dbias"] = be.download(dbias)
    return r


def check_block(be, shape, act=1, use_res=True, use_conv=True, reference_init=False, tol=TOL, spec_pack=False):
    d = block_inputs(shape, reference_init=reference_init)
    r = run_block(be, shape, d, act=act, use_res=use_res, use_conv=use_conv, spec_pack=spec_pack)
    actn = "gelu" if act else None
    wc = d["wc"] if use_conv else None
    bias = d["bias"] if use_conv else None
    res = d["res"] if use_res else None
    out, pre, X = so.fno_block_forward(d["h"], d["vb"], d["w1"], d["w2"], wc, bias, res, actn)
    ref = so.fno_block_backward(d["h"], d["vb"], d["w1"], d["w2"], wc, bias, res, actn, d["g"])
    errs = dict(out=so.rel_l2(r["out"], out), X=so.rel_l2(r["X"], X), dh=so.rel_l2(r["dh"], ref["dh"]),
                dw1=so.rel_l2(r["dw1"], ref["dw1"]), dw2=so.rel_l2(r["dw2"], ref["dw2"]))
    if act:
        errs["pre"] = so.rel_l2(r["pre"], pre)
        if use_res:
            errs["dres"] = so.rel_l2(r["dres"], ref["dres"])
    if use_conv:
        errs["dwc"] = so.rel_l2(r["dwc"], ref["dwc"])
        errs["dbias"] = so.rel_l2(r["dbias"], ref["dbias"])
    for k, v in errs.items():
        assert v < tol, f"block {k} {shape} act={act} res={use_res} conv={use_conv}: rel L2 {v:.3e}"
    return errs


def check_tc_pack(be, K=21, N=12, lda=14):
    """pdes_gemm_tc_pack: hi/lo TF32 split written in the UMMA canonical K-major layout (8 rows x 16 bytes core
    matrices, no swizzle), one (hi, lo) block pair per 16-channel chunk."""
    rng = _rng(11)
    Wt = rng.standard_normal((K, lda)).astype(np.float32)
    n_f = be.lib.pdes_gemm_tc_pack_floats(K, N)
    npad, nch = (N + 15) // 16 * 16, (K + 15) // 16
    assert n_f == nch * 2 * npad * 16
    out = be.empty((n_f,))
    dWt = be.upload(Wt)                       # keep the buffer alive across the call
    be.check(be.lib.pdes_gemm_tc_pack(be.ptr(dWt), lda, K, N, be.ptr(out), be.stream))
    got = be.download(out).reshape(nch, 2, npad * 16)
    hi_ref = (Wt.view(np.uint32) & np.uint32(0xffffe000)).view(np.float32)
    lbo = (npad // 8) * 32
    for c in range(nch):
        for kk in range(16):
            for n in range(npad):
                k = c * 16 + kk
                off = (kk // 4) * lbo + (n // 8) * 32 + (n % 8) * 4 + (kk % 4)
                w = Wt[k, n] if (k < K and n < N) else np.float32(0)
                h = hi_ref[k, n] if (k < K and n < N) else np.float32(0)
                assert got[c, 0, off] == h and got[c, 1, off] == np.float32(w - h), (c, kk, n)
    # hi + lo reconstructs the weight exactly and lo is below 2^-10 |w|
    assert np.all(np.abs(got[:, 1]) <= np.abs(got[:, 0]) * 2.0 ** -10 + 1e-30)


def check_inverse_tc(be, shape, act=1, backward_scale=0, with_spectral=True):
    """tcgen05 K3b (3xTF32) against the float64 oracle; same inputs as check_inverse."""
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(3)
    nsplit = 2
    P = (rng.standard_normal((nsplit, B, Cout, 2 * m1, m2)) + 1j * rng.standard_normal((nsplit, B, Cout, 2 * m1, m2))).astype(np.complex64)
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    wc = (rng.standard_normal((Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    bias = rng.standard_normal(Cout).astype(np.float32)
    res = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    tab = be.tables(H, W, m1, m2)
    dP = be.upload(P)
    Z = be.empty((B, H, 2 * m2, Cout))
    be.check(be.lib.pdes_inv_h(be.ptr(dP), nsplit, B, Cout, H, m1, m2, be.ptr(tab), be.ptr(Z), be.stream))
    wct = be.upload(np.ascontiguousarray(wc.T))
    pack = be.empty((be.lib.pdes_gemm_tc_pack_floats(Cin, Cout),))
    be.check(be.lib.pdes_gemm_tc_pack(be.ptr(wct), Cout, Cin, Cout, be.ptr(pack), be.stream))
    dx0, dx1 = be.upload(x0), (be.upload(x1) if C1 else None)
    if not be.lib.pdes_inv_w_gemm_tc_ok(Cout, Cin, H, W, m2 if with_spectral else 0, be.ptr(dx0), be.ptr(dx1)):
        return None                     # shape stays on the FFMA kernel (odd H*W, or too many rows per 128-pixel tile)
    dbias, dres = be.upload(bias), be.upload(res)
    out = be.empty((B, Cout, H, W))
    pre = be.empty((B, Cout, H, W))
    be.check(be.lib.pdes_inv_w_gemm_tc(be.ptr(Z) if with_spectral else None, be.ptr(pack), be.ptr(dx0), C0, be.ptr(dx1), C1,
                                       be.ptr(dbias), be.ptr(dres), be.ptr(tab), backward_scale, be.ptr(out), be.ptr(pre),
                                       B, Cout, H, W, m1, m2, act, be.stream))
    xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    ref = np.einsum("oi,bihw->bohw", wc.astype(np.float64), xin) + bias[None, :, None, None] + res
    if with_spectral:
        lscale = np.ones(m2) if backward_scale else so.hermitian_scale(H, W, m2)
        ref = ref + so.inv_pruned(P.astype(np.complex128).sum(axis=0), H, W, lscale)
    e_pre = so.rel_l2(be.download(pre), ref)
    e_out = so.rel_l2(be.download(out), so.gelu(ref) if act == 1 else ref)
    assert e_pre < TOL and e_out < TOL, f"tc inverse {shape}: pre {e_pre:.3e} out {e_out:.3e}"
    return e_pre, e_out


def check_groupnorm(be, B=3, C=12, HW=35, G=4, act=1):
    """Fused GroupNorm + GELU forward/backward against a float64 numpy restatement."""
    rng = _rng(21)
    x = (rng.standard_normal((B, C, HW)) * 1.7 + 0.4).astype(np.float32)
    dy = rng.standard_normal((B, C, HW)).astype(np.float32)
    gamma = (1 + 0.3 * rng.standard_normal(C)).astype(np.float32)
    beta = (0.2 * rng.standard_normal(C)).astype(np.float32)
    eps = 1e-5
    import ctypes
    dx_, dy_, dg_, db_ = be.upload(x), be.upload(dy), be.upload(gamma), be.upload(beta)
    y = be.empty((B, C, HW)); stats = be.empty((B * G * 2,))
    nws = (be.lib.pdes_gn_workspace_bytes(B, C, HW, G) + 3) // 4
    ws = be.empty((nws + 2,))
    wsp = (be.ptr(ws) + 7) // 8 * 8
    be.check(be.lib.pdes_gn_act_forward(be.ptr(dx_), be.ptr(dg_), be.ptr(db_), ctypes.c_float(eps), be.ptr(y), be.ptr(stats),
                                        wsp, B, C, HW, G, act, be.stream))
    gx = be.empty((B, C, HW)); gg = be.empty((C,)); gb = be.empty((C,))
    be.check(be.lib.pdes_gn_act_backward(be.ptr(dy_), be.ptr(dx_), be.ptr(dg_), be.ptr(db_), be.ptr(stats), be.ptr(gx),
                                         be.ptr(gg), be.ptr(gb), wsp, B, C, HW, G, act, be.stream))
    xd = x.astype(np.float64).reshape(B, G, -1)
    mean = xd.mean(axis=2, keepdims=True); var = xd.var(axis=2, keepdims=True)
    rstd = 1.0 / np.sqrt(var + eps)
    xh = ((xd - mean) * rstd).reshape(B, C, HW)
    z = xh * gamma[None, :, None] + beta[None, :, None]
    yr = so.gelu(z) if act else z
    dz = dy.astype(np.float64) * (so.gelu_grad(z) if act else 1.0)
    dgam = (dz * xh).sum(axis=(0, 2)); dbet = dz.sum(axis=(0, 2))
    gdz = (dz * gamma[None, :, None]).reshape(B, G, -1)
    xhg = xh.reshape(B, G, -1)
    m1 = gdz.mean(axis=2, keepdims=True); m2 = (gdz * xhg).mean(axis=2, keepdims=True)
    dxr = (rstd * (gdz - m1 - xhg * m2)).reshape(B, C, HW)
    errs = dict(y=so.rel_l2(be.download(y), yr), dx=so.rel_l2(be.download(gx), dxr),
                dgamma=so.rel_l2(be.download(gg), dgam), dbeta=so.rel_l2(be.download(gb), dbet))
    for k, v in errs.items():
        assert v < TOL, f"groupnorm {k}: {v:.3e}"
    return errs


def check_wgrad_tc(be, shape):
    """tcgen05 weight/bias gradient against the float64 oracle (returns None when the shape stays on the FFMA kernel)."""
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    rng = _rng(4)
    g = rng.standard_normal((B, Cout, H, W)).astype(np.float32)
    x0 = rng.standard_normal((B, C0, H, W)).astype(np.float32)
    x1 = rng.standard_normal((B, C1, H, W)).astype(np.float32) if C1 else None
    dg, dx0, dx1 = be.upload(g), be.upload(x0), (be.upload(x1) if C1 else None)
    nws = be.lib.pdes_wgrad_tc_workspace_floats(Cout, Cin)
    if nws == 0:
        return None
    ws = be.empty((nws,))
    dW = be.empty((Cout, Cin)); db = be.empty((Cout,))
    rc = be.lib.pdes_wgrad_tc(be.ptr(dg), be.ptr(dx0), C0, be.ptr(dx1), C1, be.ptr(dW), be.ptr(db), be.ptr(ws), B, Cout, H * W, be.stream)
    if rc == 2:
        return None
    be.check(rc)
    xin = x0 if x1 is None else np.concatenate([x0, x1], axis=1)
    e2 = so.rel_l2(be.download(dW), np.einsum("bohw,bihw->oi", g.astype(np.float64), xin))
    e3 = so.rel_l2(be.download(db), g.astype(np.float64).sum(axis=(0, 2, 3)))
    assert e2 < TOL and e3 < TOL, f"wgrad_tc {shape}: dW {e2:.3e} dbias {e3:.3e}"
    return e2, e3


def check_timeconv(be, B=2, HW=200, act=1, tw=25):
    """Fused temporal decoder (TimeConvDense Conv1d stack, dec_grid.py:97-146) forward/backward vs float64 numpy."""
    rng = _rng(33)
    KA, KB = 13, 8
    L1 = (3 * tw - KA) // 2 + 1
    z = rng.standard_normal((B, 3 * tw, HW)).astype(np.float32)
    gy = rng.standard_normal((B, tw, HW)).astype(np.float32)
    w1 = (0.3 * rng.standard_normal((2, 1, KA))).astype(np.float32); b1 = (0.1 * rng.standard_normal(2)).astype(np.float32)
    w2 = (0.3 * rng.standard_normal((1, 2, KB))).astype(np.float32); b2 = (0.1 * rng.standard_normal(1)).astype(np.float32)
    dz_, dg_, dw1, db1, dw2, db2 = (be.upload(a) for a in (z, gy, w1, b1, w2, b2))
    out = be.empty((B, tw, HW))
    be.check(be.lib.pdes_timeconv_forward(be.ptr(dz_), be.ptr(dw1), be.ptr(db1), be.ptr(dw2), be.ptr(db2), be.ptr(out), B, HW, tw,
                                          act, be.stream))
    gz = be.empty((B, 3 * tw, HW)); gw1 = be.empty((2, 1, KA)); gb1 = be.empty((2,)); gw2 = be.empty((1, 2, KB)); gb2 = be.empty((1,))
    ws = be.empty((max(1, be.lib.pdes_timeconv_bwd_workspace_floats(B, HW, tw)),))
    be.check(be.lib.pdes_timeconv_backward(be.ptr(dz_), be.ptr(dg_), be.ptr(dw1), be.ptr(db1), be.ptr(dw2), be.ptr(db2), be.ptr(gz),
                                           be.ptr(gw1), be.ptr(gb1), be.ptr(gw2), be.ptr(gb2), be.ptr(ws), B, HW, tw, act, be.stream))
    # float64 restatement
    zd, gd = z.astype(np.float64), gy.astype(np.float64)
    idx = 2 * np.arange(L1)[:, None] + np.arange(KA)[None, :]                        # [L1][KA]
    zw = zd[:, idx, :]                                                                # [B][L1][KA][HW]
    h = np.einsum("ok,blkp->bolp", w1[:, 0].astype(np.float64), zw) + b1[None, :, None, None]
    a = so.gelu(h) if act else h
    tdx = np.arange(tw)[:, None] + np.arange(KB)[None, :]                             # [tw][KB]
    aw = a[:, :, tdx, :]                                                              # [B][2][tw][KB][HW]
    y = np.einsum("ok,botkp->btp", w2[0].astype(np.float64), aw) + b2[0]
    da = np.zeros_like(a)
    for k in range(KB):
        da[:, :, k:k + tw, :] += w2[0][None, :, k, None, None] * gd[:, None, :, :]
    dh = da * (so.gelu_grad(h) if act else 1.0)
    dzr = np.zeros_like(zd)
    for k in range(KA):
        dzr[:, k:k + 2 * L1:2, :] += np.einsum("o,bolp->blp", w1[:, 0, k].astype(np.float64), dh)
    dw1r = np.einsum("bolp,blkp->ok", dh, zw)[:, None, :]
    db1r = dh.sum(axis=(0, 2, 3))
    dw2r = np.einsum("btp,botkp->ok", gd, aw)[None]
    db2r = gd.sum()[None]
    errs = dict(y=so.rel_l2(be.download(out), y), dz=so.rel_l2(be.download(gz), dzr), dw1=so.rel_l2(be.download(gw1), dw1r),
                db1=so.rel_l2(be.download(gb1), db1r), dw2=so.rel_l2(be.download(gw2), dw2r), db2=so.rel_l2(be.download(gb2), db2r))
    for k, v in errs.items():
        assert v < TOL, f"timeconv {k}: {v:.3e}"
    return errs


def check_constrain(be, B=3, tw=5, H=9, W=7, n_spatial=2, use_tanh=1, use_mask=1, use_volume=1):
    """Fused output constraints (csrc/constrain.cu) vs the reference formulas of activation_wrapper.py:33-106 and
    dec_grid.py:8-23 evaluated in float64."""
    rng = _rng(13)
    HW = H * W
    x = (rng.random((B, 1, tw, H, W)) * 0.5 + 0.1).astype(np.float32)
    delta = rng.standard_normal((B, 1, tw, H, W)).astype(np.float32)
    mask = (rng.random((B, n_spatial, H, W)) < 0.2).astype(np.float32)
    dt, pct = np.float32(0.01), np.float32(1 / 25)
    steps = np.cumsum(np.full(tw, dt, dtype=np.float32), dtype=np.float32)
    cap = np.cumsum(np.full(tw, pct, dtype=np.float32), dtype=np.float32)
    out = be.empty((B, 1, tw, H, W))
    d_delta, d_x, d_mask, d_steps, d_cap = (be.upload(a) for a in (delta, x, mask, steps, cap))     # keep the buffers alive
    be.check(be.lib.pdes_constrain_forward(be.ptr(d_delta), be.ptr(d_x), be.ptr(d_mask), n_spatial * HW, be.ptr(d_steps),
                                           be.ptr(d_cap), be.ptr(out), B, tw, HW, use_tanh, use_mask, use_volume, be.stream))
    xd, dd, m = x.astype(np.float64), delta.astype(np.float64), mask[:, 0].astype(np.float64)[:, None, None]
    u = xd[:, :, -1:] + steps.astype(np.float64)[None, None, :, None, None] * dd
    if use_tanh:
        u = np.tanh(u)
    if use_mask:
        u = u - m * u
    if use_volume:
        new = u.sum(axis=(3, 4))
        prev = np.broadcast_to(xd[:, :, -1].sum(axis=(2, 3))[:, :, None], new.shape)
        c = cap.astype(np.float64)[None, None, :]
        dif = np.tanh((1 - new / prev) * 100 / c) / 100 * c
        u = u / new[..., None, None] * ((1 - dif) * prev)[..., None, None]
        if use_mask:
            u = u - m * u
    err = so.rel_l2(be.download(out), u)
    assert err < TOL, f"constrain_forward tanh={use_tanh} mask={use_mask} volume={use_volume}: rel L2 {err:.3e}"
    return err
