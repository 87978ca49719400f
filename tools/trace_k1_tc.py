"""In-kernel clock trace of CTA 0 of the tcgen05 K1 (library built with PDES_NVCC_EXTRA=-DPDES_K1_TRACE)."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_pde_surrogates_b200 import _native  # noqa: E402

lib = _native.library()
raw = ctypes.CDLL(os.path.join(ROOT, "neural_pde_surrogates_b200", "lib", "libpdes_b200.so"))
dev = torch.device("cuda:0")
B, C0, C1, H, W, m1, m2 = 16, 192, 1, 96, 64, 10, 10
n = lib.pdes_tables_floats(H, W, m1, m2)
buf = np.zeros(n, dtype=np.float32)
_native.check(lib, lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data))
tab = torch.from_numpy(buf).to(dev)
x0, x1 = torch.randn(B, C0, H, W, device=dev), torch.randn(B, C1, H, W, device=dev)
X = torch.empty(B, C0 + C1, 2 * m1, m2, dtype=torch.complex64, device=dev)
X2 = torch.zeros(lib.pdes_mix_tc_x2_floats(B, C0 + C1, m1, m2), device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    _native.check(lib, lib.pdes_dft_fwd2(x0.data_ptr(), C0, x1.data_ptr(), C1, B, H, W, m1, m2, tab.data_ptr(), 0, X.data_ptr(), X2.data_ptr(), st))
torch.cuda.synchronize()
out = np.zeros(5 * 16 * 4, dtype=np.int64)
raw.pdes_debug_k1_trace(ctypes.c_void_p(out.ctypes.data))
t = out.reshape(5, 16, 4)
t0 = t[t > 0].min()
names = ["tma  (top, after empty wait)", "mma  (top, after waits, stage1 committed, stage2 issued)", "lo   (top, full, d1_full(it-2), arrived)",
         "a2   (top, d1_full, d2_full(it-2), arrived)", "out  (top, d2_full, stored)"]
for r in range(5):
    print(names[r])
    for it in range(12):
        print("   it", it, [int(v - t0) if v > 0 else None for v in t[r, it]])
