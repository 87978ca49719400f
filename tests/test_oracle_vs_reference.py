"""Live pin of the oracle against the UNMODIFIED reference modules (CPU).  Runs wherever the reference sources can be
imported -- /root/reference/src in the build container, or the verbatim copy `oracle/make_ref.sh` leaves in
oracle/_ref/src -- and is skipped otherwise (the committed fixtures in tests/golden/ then carry the pin, see
test_oracle_golden.py).  Unlike the fixtures, the inputs here are drawn fresh: shapes with odd sizes, 2*m1 > H (the
reference's second weight block overwrites the first, proc_fno.py:266-269) and m2 = W/2 + 1 (Nyquist column kept).
Tolerance: fp32 reference vs fp64 oracle, rel L2 <= 1e-5 (north_star)."""
import numpy as np
import pytest
import torch

from oracle import reference_loader as rl
from oracle import spectral_oracle as so

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="reference sources not importable here")

TOL = 1e-5
# (B, Cin, Cout, H, W, m1, m2)
CASES = [(2, 3, 4, 12, 8, 3, 4), (1, 5, 2, 9, 7, 2, 3), (2, 4, 4, 6, 10, 5, 6), (1, 2, 3, 16, 16, 8, 9), (3, 6, 5, 20, 12, 4, 5)]


@pytest.fixture(scope="module")
def ref():
    return rl.load_reference()


@pytest.mark.parametrize("case", CASES)
def test_spectral_conv2d_forward_backward(ref, case):
    """SpectralConv2d.forward (proc_fno.py:257-288) and its autograd gradients against the dense-DFT restatement."""
    B, Cin, Cout, H, W, m1, m2 = case
    torch.manual_seed(sum(case))
    layer = ref.proc_fno.SpectralConv2d(Cin, Cout, (m1, m2))
    with torch.no_grad():                                # the reference initialises with a 1/(Cin*Cout) scale: make the
        layer.weights1.mul_(Cin * Cout)                  # output O(1) so that the relative error is meaningful
        layer.weights2.mul_(Cin * Cout)
    x = torch.randn(B, Cin, H, W, requires_grad=True)
    g = torch.randn(B, Cout, H, W)
    y = layer(x)
    y.backward(g)
    w1, w2 = layer.weights1.detach().numpy(), layer.weights2.detach().numpy()
    yo, X = so.spectral_conv2d_forward(x.detach().numpy(), w1, w2)
    gx, gw1, gw2 = so.spectral_conv2d_backward(x.detach().numpy(), w1, w2, g.numpy(), X)
    assert so.rel_l2(yo, y.detach().numpy()) < TOL
    assert so.rel_l2(gx, x.grad.numpy()) < TOL
    live = so.live_rows(H, m1)[:m1] > 0               # rows of weights1 the reference overwrites get a zero gradient
    assert so.rel_l2(gw1, layer.weights1.grad.numpy()) < TOL
    assert so.rel_l2(gw2, layer.weights2.grad.numpy()) < TOL
    if not live.all():
        assert np.all(layer.weights1.grad.numpy()[:, :, ~live] == 0)


@pytest.mark.parametrize("case", CASES[:3])
def test_fno_layer_forward(ref, case):
    """FNO_Layer.forward = act(conv(x) + w(x)) (proc_fno.py:133-155) against fno_block_forward."""
    B, C, _, H, W, m1, m2 = case
    torch.manual_seed(7 + sum(case))
    layer = ref.proc_fno.FNO_Layer(C, num_spatial_dims=2, kernel_size=1, modes=(m1, m2))
    x = torch.randn(B, C, H, W)
    with torch.no_grad():
        y = layer(x)
    conv = layer.conv
    out = so.fno_block_forward(x.numpy(), None, conv.weights1.detach().numpy(), conv.weights2.detach().numpy(),
                               layer.w.weight.detach().numpy().reshape(C, C), layer.w.bias.detach().numpy(), None, "gelu")
    assert so.rel_l2(out[0], y.numpy()) < TOL
