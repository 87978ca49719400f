"""Skeleton cost of the K3b kernel vs number of chunks / tiles (trace build)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import _native
lib = _native.library()
dev = torch.device("cuda:0")
H, W, m1, m2 = 96, 64, 10, 10
st = torch.cuda.current_stream().cuda_stream
p = lambda t: None if t is None else t.data_ptr()
def run(B, C0, Cout, act):
    h = torch.randn(B, C0, H, W, device=dev)
    wct = torch.randn(C0, Cout, device=dev)
    out = torch.empty(B, Cout, H, W, device=dev)
    pack = torch.empty(lib.pdes_gemm_tc_pack_floats(C0, Cout), device=dev)
    lib.pdes_gemm_tc_pack(p(wct), Cout, C0, Cout, p(pack), st)
    fn = lambda: lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, None, 0, None, None, None, 0, p(out), None, B, Cout, H, W, m1, m2, act, st)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3
for B, C0, Cout in ((16, 192, 192), (16, 96, 192), (16, 48, 192), (16, 16, 192), (8, 192, 192), (4, 192, 192), (1, 192, 192), (16, 192, 64)):
    tiles = B * H * W // 128
    print(f"B={B:2d} C0={C0:3d} N={Cout:3d} tiles/CTA={tiles/148:4.1f} chunks={C0//16:2d}: skeleton {run(B, C0, Cout, 77 + 256*15):6.1f} us  "
          f"no-epilogue {run(B, C0, Cout, 77):6.1f} us  store-only {run(B, C0, Cout, 0):6.1f} us")
