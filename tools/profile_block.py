"""Tiny driver for ncu: a few forward+backward passes of one fused U-FNO block at the config shape (B from argv)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import neural_pde_surrogates_b200 as npb  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = npb.FNO_Layer(hidden_dim=193, hidden_dim_out=192, num_spatial_dims=2, modes=10, activation=None).to(dev)
h = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
vb = (torch.rand(B, 1, 96, 64, device=dev) < 0.1).float()
res = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
g = torch.randn(B, 192, 96, 64, device=dev)
act = torch.nn.GELU()
for _ in range(iters):
    out = layer.fused(h, vb, res, act)
    out.backward(g)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
