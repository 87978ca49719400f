"""One forward + backward of a full-width U-FNO block tail and one opt-in tcgen05 K1 call (quick driver for profilers)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_pde_surrogates_b200 as npb  # noqa: E402

B = int(os.environ.get("B", "4"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = npb.FNO_Layer(hidden_dim=193, hidden_dim_out=192, num_spatial_dims=2, modes=10, activation=None).to(dev)
h = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
vb = (torch.rand(B, 1, 96, 64, device=dev) < 0.1).float()
res = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
out = layer.fused(h, vb, res, torch.nn.GELU())
out.sum().backward()
with torch.no_grad():
    out2 = layer.fused(h, vb, res, torch.nn.GELU())
os.environ["PDES_K1_TC"] = "1"
with torch.no_grad():
    out3 = layer.fused(h, vb, res, torch.nn.GELU())
torch.cuda.synchronize()
print("ok", float(out.detach().abs().mean()), float((out3 - out2).abs().max()), float(h.grad.abs().mean()))
