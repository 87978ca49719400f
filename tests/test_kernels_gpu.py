"""-m gpu parity tests: the real sm_100a library, called through the C ABI, against the float64 oracle.
Same cases as the CPU emulation suite plus the reference's full config shapes (cfg_twophase_ufno.py)."""
import pytest

import kernel_cases as kc

pytestmark = pytest.mark.gpu

# cfg_twophase_ufno.py: width 192, n_cond 1 (mask), grid 96x64, modes 10x10
FULL_SHAPES = [
    (4, 192, 1, 192, 96, 64, 10, 10),
    (2, 128, 1, 128, 64, 64, 16, 16),
    (16, 192, 1, 192, 96, 64, 10, 10),       # the bench batch at full channels (B > 8 tile configs of the mix kernels)
]


@pytest.fixture(scope="module")
def be():
    from backends import CudaBackend
    return CudaBackend()


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
@pytest.mark.parametrize("herm", [0, 1])
def test_dft_fwd(be, shape, herm):
    kc.check_dft_fwd(be, shape, herm)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_mix(be, shape):
    kc.check_mix(be, shape)


# larger batches reach every tile configuration of the TMA-fed mix kernels (B > 8 / 5..8 / <= 4, sample tiles with
# out-of-bounds rows, several mode tiles per weight block, dead rows) and the TMA weight-gradient kernel (B >= 8)
@pytest.mark.parametrize("shape", [(9, 5, 1, 12, 12, 8, 3, 4), (16, 24, 1, 20, 16, 16, 4, 5), (6, 7, 0, 9, 8, 8, 2, 3),
                                   (8, 4, 0, 6, 8, 8, 5, 2), (8, 4, 1, 6, 32, 64, 16, 16), (17, 9, 0, 17, 96, 64, 10, 10)])
def test_mix_tma_configs(be, shape):
    kc.check_mix(be, shape)


# K2 on tcgen05: small edge-case shapes (Nyquist column, dead rows, odd sizes), two output-channel tiles (Cout 140),
# B > 8 (N = 32), B = 32 (N = 64, the widest accumulator), and the shipped shape at the bench batch
@pytest.mark.parametrize("shape", kc.SMALL_SHAPES + [(3, 20, 1, 140, 12, 8, 3, 4), (16, 24, 1, 20, 16, 16, 4, 5),
                                                     (32, 17, 0, 9, 8, 8, 2, 3), (4, 192, 1, 192, 96, 64, 10, 10),
                                                     (16, 192, 1, 192, 96, 64, 10, 10)])
def test_spectral_tc(be, shape):
    kc.check_spectral_tc(be, shape, run_mma=True)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES + FULL_SHAPES)
def test_block_fwd_bwd_packed_spectral_weights(be, shape):
    kc.check_block(be, shape, act=1, use_res=True, use_conv=True, spec_pack=True)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_inverse_full(be, shape):
    kc.check_inverse(be, shape, with_gemm=True, act=1)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES[:3])
def test_inverse_spectral_only(be, shape):
    kc.check_inverse(be, shape, with_gemm=False, act=0, backward_scale=1)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_pointwise(be, shape):
    kc.check_pointwise(be, shape)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES + FULL_SHAPES)
def test_block_fwd_bwd(be, shape):
    kc.check_block(be, shape, act=1, use_res=True, use_conv=True)


def test_block_variants(be):
    kc.check_block(be, kc.SMALL_SHAPES[2], act=0, use_res=False, use_conv=True)
    kc.check_block(be, kc.SMALL_SHAPES[1], act=1, use_res=False, use_conv=True)
    kc.check_block(be, kc.SMALL_SHAPES[0], act=0, use_res=False, use_conv=False)


def test_block_reference_init_full_shape(be):
    """Full config shape with the reference's own weight init scale (1/(Cin*Cout) * U[0,1), proc_fno.py:239-243)."""
    kc.check_block(be, FULL_SHAPES[0], reference_init=True)


def test_large_grid(be):
    """BASELINE config 5 shape class: 256x256 grid, 32x32 modes (exercises the chunked shared-memory paths)."""
    kc.check_block(be, (1, 16, 1, 16, 256, 256, 32, 32))


# ---- tcgen05 / TMEM path (3xTF32) ---------------------------------------------------------------------------
@pytest.fixture(params=[2, 1, 0], ids=["tc_v3", "tc_v2", "ffma"])
def tc_mode(request, be):
    prev = be.lib.pdes_get_tensor_core_mode()
    be.lib.pdes_set_tensor_core_mode(request.param)
    yield request.param
    be.lib.pdes_set_tensor_core_mode(prev)        # the process default (2) must survive for the tests that follow


def test_tc_weight_pack_layout(be):
    kc.check_tc_pack(be)
    kc.check_tc_pack(be, K=193, N=192, lda=192)


TC_SHAPES = kc.SMALL_SHAPES + FULL_SHAPES + [(2, 20, 1, 24, 10, 24, 3, 4), (1, 40, 0, 200, 16, 32, 5, 7)]


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_inverse_gemm(be, shape):
    kc.check_inverse_tc(be, shape)


def test_tc_path_is_taken_for_config_shapes(be):
    for shape in FULL_SHAPES + TC_SHAPES[-2:]:
        assert kc.check_inverse_tc(be, shape) is not None


def test_tc_gemm_only_and_backward_scale(be):
    kc.check_inverse_tc(be, kc.SMALL_SHAPES[2], act=0, with_spectral=False)
    kc.check_inverse_tc(be, kc.SMALL_SHAPES[5], act=0, backward_scale=1)


@pytest.mark.parametrize("shape", [kc.SMALL_SHAPES[3], kc.SMALL_SHAPES[6], FULL_SHAPES[0]])
def test_block_both_modes(be, shape, tc_mode):
    assert be.lib.pdes_get_tensor_core_mode() == tc_mode
    kc.check_block(be, shape, act=1, use_res=True, use_conv=True)


def test_groupnorm_act(be):
    kc.check_groupnorm(be, B=3, C=12, HW=35, G=4, act=1)
    kc.check_groupnorm(be, B=2, C=193, HW=96 * 64, G=1, act=1)     # the U-Net residual-block shape
    kc.check_groupnorm(be, B=2, C=192, HW=100 * 68, G=8, act=1)    # the final norm of the U-Net


@pytest.mark.parametrize("shape", TC_SHAPES)
def test_tc_wgrad(be, shape):
    kc.check_wgrad_tc(be, shape)


def test_tc_wgrad_path_is_taken_for_config_shapes(be):
    for shape in FULL_SHAPES + [(2, 20, 1, 24, 8, 24, 3, 4), (1, 40, 0, 200, 16, 32, 5, 7)]:
        assert kc.check_wgrad_tc(be, shape) is not None


def test_timeconv_decoder(be):
    kc.check_timeconv(be, B=2, HW=150, act=1)
    kc.check_timeconv(be, B=1, HW=128, act=0)
    kc.check_timeconv(be, B=4, HW=96 * 64, act=1)     # the decoder shape of the twophase config


def test_fused_output_constraints(be):
    kc.check_constrain(be)
    kc.check_constrain(be, B=4, tw=25, H=96, W=64, n_spatial=1)
    kc.check_constrain(be, B=2, tw=25, H=12, W=8, n_spatial=1, use_volume=0)
    kc.check_constrain(be, B=1, tw=3, H=20, W=20, use_tanh=0, use_mask=0)


def test_tc_inverse_gemm_run_to_run_bitwise(be):
    """The bench-shape K3b launch repeated 150 times must be bit-identical.  Before the raw-ring release was moved behind
    the consumption of the loaded rows, ~9 % of B = 16 launches had one 32-pixel quadrant of one tile computed with a
    k-row of the chunk 8 ahead (tools/stress_k3b.py); a parity check of a single launch misses that most of the time."""
    import torch
    lib, p, ck = be.lib, be.ptr, be.check
    B, C0, C1, Cout, H, W, m1, m2 = FULL_SHAPES[2]
    Cin = C0 + C1
    g = torch.Generator(device=be.dev).manual_seed(5)
    rn = lambda *s: torch.randn(*s, device=be.dev, generator=g)
    Z, x0, x1 = rn(B, H, 2 * m2, Cout), rn(B, C0, H, W), rn(B, C1, H, W)
    wc, bias, res = rn(Cout, Cin) / Cin ** 0.5, rn(Cout), rn(B, Cout, H, W)
    tab = be.tables(H, W, m1, m2)
    pack = be.empty((lib.pdes_gemm_tc_pack_floats(Cin, Cout),))
    ck(lib.pdes_gemm_tc_pack_t(p(wc), Cin, Cin, Cout, p(pack), be.stream))
    flush = torch.zeros(48 * 1024 * 1024, device=be.dev)

    def run():
        out = be.empty((B, Cout, H, W))
        ck(lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(x0), C0, p(x1), C1, p(bias), p(res), p(tab), 0, p(out), None,
                                  B, Cout, H, W, m1, m2, 1, be.stream))
        return out

    first = run()
    for i in range(150):
        if i % 2:
            flush.add_(1.0)                       # cold and warm L2: the race window depended on the copy latency
        assert torch.equal(run().view(torch.int32), first.view(torch.int32)), f"launch {i} differs from launch 0"


@pytest.mark.parametrize("B", [4, 16])
def test_block_forward_backward_run_to_run_bitwise(be, B):
    """No kernel of the block uses atomics, so every output of pdes_block_forward / pdes_block_backward must repeat bit for
    bit.  100 launches with alternating cold / warm L2 at the shipped shape: the check that exposes ring races (a
    consumer releasing a stage before its loads landed, a producer overtaking a reader), which corrupt one tile in
    thousands and slip through a single parity comparison."""
    import torch
    lib, p, ck = be.lib, be.ptr, be.check
    _, C0, C1, Cout, H, W, m1, m2 = FULL_SHAPES[2]
    Cin, st = C0 + C1, be.stream
    g = torch.Generator(device=be.dev).manual_seed(11)
    rn = lambda *s: torch.randn(*s, device=be.dev, generator=g)
    h, vb, res, gout = rn(B, C0, H, W), rn(B, C1, H, W), rn(B, Cout, H, W), rn(B, Cout, H, W)
    w1 = torch.view_as_complex(rn(Cin, Cout, m1, m2, 2).contiguous()) / Cin
    w2 = torch.view_as_complex(rn(Cin, Cout, m1, m2, 2).contiguous()) / Cin
    wc, bias = rn(Cout, Cin) / Cin ** 0.5, rn(Cout)
    tab = be.tables(H, W, m1, m2)
    wsp = be.empty((lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2),))
    ck(lib.pdes_mix_tc_pack(p(w1), p(w2), p(wsp), Cin, Cout, H, m1, m2, st))
    pack_f = be.empty((lib.pdes_gemm_tc_pack_floats(Cin, Cout),))
    ck(lib.pdes_gemm_tc_pack_t(p(wc), Cin, Cin, Cout, p(pack_f), st))
    ws = be.empty((lib.pdes_block_fwd_workspace_floats(B, Cin, Cout, H, W, m1, m2),))
    wsb = be.empty((lib.pdes_block_bwd_workspace_floats(B, C0, C1, Cout, H, W, m1, m2),))
    flush = torch.zeros(48 * 1024 * 1024, device=be.dev)

    def run():
        X = be.empty((B, Cin, 2 * m1, m2), complex_=True)
        out, pre = be.empty((B, Cout, H, W)), be.empty((B, Cout, H, W))
        ck(lib.pdes_block_forward(p(h), C0, p(vb), C1, p(w1), p(w2), p(wsp), p(wc), p(pack_f), p(bias), p(res), p(tab), p(X),
                                  p(ws), p(out), p(pre), B, Cout, H, W, m1, m2, 1, st))
        gpre, dh = be.empty((B, Cout, H, W)), be.empty((B, C0, H, W))
        gw1, gw2 = be.empty(tuple(w1.shape), complex_=True), be.empty(tuple(w2.shape), complex_=True)
        dwc, dbias = be.empty((Cout, Cin)), be.empty((Cout,))
        ck(lib.pdes_block_backward(p(gout), p(pre), p(h), C0, p(vb), C1, p(X), p(w1), p(w2), p(wsp), p(wc), None, p(tab), p(wsb),
                                   p(gpre), p(dh), p(gw1), p(gw2), p(dwc), p(dbias), B, Cout, H, W, m1, m2, 1, st))
        return dict(X=torch.view_as_real(X), out=out, pre=pre, gpre=gpre, dh=dh, gw1=torch.view_as_real(gw1),
                    gw2=torch.view_as_real(gw2), dwc=dwc, dbias=dbias)

    first = run()
    assert all(torch.isfinite(v).all() for v in first.values())
    for i in range(100):
        if i % 2:
            flush.add_(1.0)
        now = run()
        bad = [k for k in first if not torch.equal(now[k].view(torch.int32), first[k].view(torch.int32))]
        assert not bad, f"launch {i}: {bad} differ from launch 0"


@pytest.mark.parametrize("herm", [0, 1])
def test_dft_fwd_tensor_core_opt_in(be, herm, monkeypatch):
    """The opt-in tcgen05 K1 (both DFT stages as 3xTF32 GEMMs with MN-major operands, csrc/spectral_dft_tc.cu) against the
    float64 oracle at the shipped grid, odd image count and two input tensors included."""
    monkeypatch.setenv("PDES_K1_TC", "1")
    for shape in [(1, 3, 1, 4, 96, 64, 10, 10), (2, 4, 1, 8, 96, 64, 10, 10), (3, 5, 0, 4, 32, 64, 4, 7), (4, 192, 1, 192, 96, 64, 10, 10)]:
        kc.check_dft_fwd(be, shape, herm)
    monkeypatch.delenv("PDES_K1_TC")
    kc.check_dft_fwd(be, (1, 3, 1, 4, 96, 64, 10, 10), herm)


@pytest.mark.parametrize("shape", [FULL_SHAPES[0], FULL_SHAPES[2], (3, 128, 1, 96, 32, 32, 6, 5), (16, 64, 0, 160, 16, 16, 4, 4)])
def test_mix_dx_tensor_core(be, shape):
    """pdes_mix_tc_dx (the adjoint of the channel mix read from the FORWARD pack as an MN-major operand) against the float64
    oracle, through its own entry point: GO2 from pdes_dft_fwd2-style mode-major data, output in K2's O2 layout."""
    import numpy as np
    from oracle import spectral_oracle as so
    lib, p, ck = be.lib, be.ptr, be.check
    B, C0, C1, Cout, H, W, m1, m2 = shape
    Cin = C0 + C1
    if not lib.pdes_mix_tc_dx_ok(B, Cin, Cout, C0, m1, m2):
        pytest.skip("shape outside pdes_mix_tc_dx")
    rng = np.random.default_rng(7)
    w1, w2 = kc._weights(rng, Cin, Cout, m1, m2)
    GO = (rng.standard_normal((B, Cout, 2 * m1, m2)) + 1j * rng.standard_normal((B, Cout, 2 * m1, m2))).astype(np.complex64)
    CoutP = (Cout + 15) // 16 * 16
    go2 = np.full((2 * m1 * m2, B, CoutP), np.nan + 1j * np.nan, dtype=np.complex64)       # pad columns must never be used
    go2[:, :, :Cout] = GO.transpose(2, 3, 0, 1).reshape(2 * m1 * m2, B, Cout)
    dw1, dw2, dgo = be.upload(w1), be.upload(w2), be.upload(go2)
    wsp = be.empty((lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2),))
    ck(lib.pdes_mix_tc_pack(p(dw1), p(dw2), p(wsp), Cin, Cout, H, m1, m2, be.stream))
    first = None
    for rep in range(3):                                              # repeated: the ring / mode hand-over must not depend on timing
        o2 = be.empty((2, 2 * m1 * m2, B, C0), complex_=True)
        ck(lib.pdes_mix_tc_dx(p(dgo), p(wsp), p(o2), B, Cin, Cout, C0, m1, m2, be.stream))
        got = be.download(o2)
        assert np.all(got[1] == 0)
        first = got if first is None else first
        assert np.array_equal(got.view(np.float32), first.view(np.float32), equal_nan=True)
    gx = first[0].reshape(2 * m1, m2, B, C0).transpose(2, 3, 0, 1)
    ref = so.mode_mix_dx(GO.astype(np.complex128), w1, w2, H)[:, :C0]
    err = so.rel_l2(gx, ref)
    assert err < kc.TOL, f"mix_tc_dx {shape}: rel L2 {err:.3e}"
