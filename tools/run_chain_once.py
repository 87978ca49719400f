"""Run the forward chain of one U-FNO block a few times at the bench shape (for ncu captures)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import neural_pde_surrogates_b200 as npb  # noqa: E402
from neural_pde_surrogates_b200 import ops  # noqa: E402

B = int(os.environ.get("B", "16"))
dev = torch.device("cuda:0")
torch.manual_seed(0)
layer = npb.FNO_Layer(hidden_dim=193, hidden_dim_out=192, num_spatial_dims=2, modes=10, activation=None).to(dev)
h = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
vb = (torch.rand(B, 1, 96, 64, device=dev) < 0.1).float()
res = torch.randn(B, 192, 96, 64, device=dev, requires_grad=True)
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for it in range(int(os.environ.get("ITERS", "4"))):
    flush.add_(1.0)
    out = layer.fused(h, vb, res, torch.nn.GELU())
    if os.environ.get("BWD", "0") == "1":
        out.sum().backward()
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
