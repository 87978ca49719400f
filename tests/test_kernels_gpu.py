"""-m gpu parity tests: the real sm_100a library, called through the C ABI, against the float64 oracle.
Same cases as the CPU emulation suite plus the reference's full config shapes (cfg_twophase_ufno.py)."""
import pytest

import kernel_cases as kc

pytestmark = pytest.mark.gpu

# cfg_twophase_ufno.py: width 192, n_cond 1 (mask), grid 96x64, modes 10x10
FULL_SHAPES = [
    (4, 192, 1, 192, 96, 64, 10, 10),
    (2, 128, 1, 128, 64, 64, 16, 16),
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
