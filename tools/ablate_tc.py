"""Ablation timing of the persistent K3b kernel (library built with PDES_NVCC_EXTRA=-DPDES_TC_TRACE): act = base + 256*mask,
mask bits: 1 no MMA, 2 no weight copy, 4 no activation copy, 8 no convert; base 77 = no epilogue, 78 = TMEM loads only."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import _native
lib = _native.library()
dev = torch.device("cuda:0")
B = 16
C0, C1, Cout, H, W, m1, m2 = 192, 1, 192, 96, 64, 10, 10
Cin = C0 + C1
n = lib.pdes_tables_floats(H, W, m1, m2)
buf = np.zeros(n, dtype=np.float32); lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data)
tab = torch.from_numpy(buf).to(dev)
h = torch.randn(B, C0, H, W, device=dev); vb = torch.randn(B, C1, H, W, device=dev)
wct = torch.randn(Cin, Cout, device=dev) / Cin ** 0.5
out = torch.empty(B, Cout, H, W, device=dev)
Z = torch.randn(B, H, 2 * m2, Cout, device=dev); res = torch.randn(B, Cout, H, W, device=dev); bias = torch.randn(Cout, device=dev)
pack = torch.empty(lib.pdes_gemm_tc_pack_floats(Cin, Cout), device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: None if t is None else t.data_ptr()
lib.pdes_gemm_tc_pack(p(wct), Cout, Cin, Cout, p(pack), st)
def run(act, spectral=True, epi=True):
    fn = lambda: lib.pdes_inv_w_gemm_tc(p(Z) if spectral else None, p(pack), p(h), C0, p(vb), C1, p(bias) if epi else None,
                                        p(res) if epi else None, p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, act, st)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5 * 1e3
for name, base in (("full epilogue", 1), ("no epilogue", 77)):
    for mask in (0, 1, 2, 4, 8, 3, 6, 7, 15):
        print(f"{name:14s} mask {mask:2d}: spectral {run(base + 256 * mask):7.1f} us   x-only {run(base + 256 * mask, spectral=False):7.1f} us")
print("epilogue-only runs (mask 15 = no MMA / copies / convert), x-only:")
print("  res+bias+GELU+store :", run(1 + 256 * 15, spectral=False))
print("  res+bias+store      :", run(0 + 256 * 15, spectral=False))
print("  store only          :", run(0 + 256 * 15, spectral=False, epi=False))
print("  GELU+store (no res) :", run(1 + 256 * 15, spectral=False, epi=False))
print("  TMEM loads only     :", run(78 + 256 * 15, spectral=False))
print("  nothing             :", run(77 + 256 * 15, spectral=False))
