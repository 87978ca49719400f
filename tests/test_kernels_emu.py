"""CPU suite: the CUDA kernel sources compiled for the CPU emulator, checked against the float64 oracle.
This validates indexing, staging, guards and the math of every kernel without a GPU; the -m gpu tests repeat the
same cases on the real sm_100a build."""
import pytest

import kernel_cases as kc
from backends import EmuBackend


@pytest.fixture(scope="module")
def be():
    return EmuBackend()


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
@pytest.mark.parametrize("herm", [0, 1])
def test_dft_fwd(be, shape, herm):
    kc.check_dft_fwd(be, shape, herm)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_mix(be, shape):
    kc.check_mix(be, shape)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_inverse_full(be, shape):
    kc.check_inverse(be, shape, with_gemm=True, act=1)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES[:3])
def test_inverse_spectral_only(be, shape):
    kc.check_inverse(be, shape, with_gemm=False, act=0, backward_scale=1)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_pointwise(be, shape):
    kc.check_pointwise(be, shape)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES)
def test_block_fwd_bwd(be, shape):
    kc.check_block(be, shape, act=1, use_res=True, use_conv=True)


@pytest.mark.parametrize("shape", kc.SMALL_SHAPES + [(3, 20, 1, 140, 12, 8, 3, 4)])
def test_spectral_tc_host_side(be, shape):
    """weight pack, K1's mode-major copy and K3a on the tcgen05 K2's output layout (the MMA kernel itself is GPU-only)."""
    kc.check_spectral_tc(be, shape, run_mma=False)


def test_block_variants(be):
    kc.check_block(be, kc.SMALL_SHAPES[2], act=0, use_res=False, use_conv=True)    # FNO layer without activation
    kc.check_block(be, kc.SMALL_SHAPES[1], act=1, use_res=False, use_conv=True)    # pure FNO layer (GELU inside)
    kc.check_block(be, kc.SMALL_SHAPES[0], act=0, use_res=False, use_conv=False)   # bare SpectralConv2d


def test_bad_arguments_raise(be):
    import numpy as np
    lib = be.lib
    x = be.upload(np.zeros((1, 1, 8, 8)))
    X = be.empty((1, 1, 4, 6), complex_=True)
    tab = be.tables(8, 8, 2, 2)
    # m2 > W//2+1 must be rejected like the reference's assert (proc_fno.py:135-139)
    with pytest.raises(ValueError):
        be.check(lib.pdes_dft_fwd(be.ptr(x), 1, None, 0, 1, 8, 8, 2, 6, be.ptr(tab), 0, be.ptr(X), be.stream))
    with pytest.raises(ValueError):
        be.check(lib.pdes_dft_fwd(None, 1, None, 0, 1, 8, 8, 2, 2, be.ptr(tab), 0, be.ptr(X), be.stream))
    with pytest.raises(ValueError):
        be.check(lib.pdes_tables_fill(8, 8, 9, 2, x.ctypes.data))


def test_tc_weight_pack_layout(be):
    kc.check_tc_pack(be)
    kc.check_tc_pack(be, K=16, N=16, lda=16)
    assert be.lib.pdes_get_tensor_core_mode() == 0          # the emulation build never takes the tcgen05 path


def test_groupnorm_act(be):
    kc.check_groupnorm(be, B=3, C=12, HW=35, G=4, act=1)       # scalar path
    kc.check_groupnorm(be, B=2, C=10, HW=64, G=1, act=1)       # GroupNorm(1, C) as in the residual blocks, vector path
    kc.check_groupnorm(be, B=2, C=8, HW=16, G=8, act=0)


def test_timeconv_decoder(be):
    kc.check_timeconv(be, B=2, HW=150, act=1)         # partial last block
    kc.check_timeconv(be, B=1, HW=128, act=0)


def test_fused_output_constraints(be):
    kc.check_constrain(be)
    kc.check_constrain(be, B=2, tw=25, H=12, W=8, n_spatial=1, use_volume=0)
    kc.check_constrain(be, B=1, tw=3, H=20, W=20, use_tanh=0, use_mask=0)
