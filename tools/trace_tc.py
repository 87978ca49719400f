"""Dump the in-kernel clock trace of the persistent tcgen05 kernel (library built with -DPDES_TC_TRACE)."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import _native
lib = _native.bind(ctypes.CDLL(os.environ['PDES_LIB'])) if os.environ.get('PDES_LIB') else _native.library()
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
ACT = int(os.environ.get('ACT', '1'))
for _ in range(3):
    lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, ACT, st)
torch.cuda.synchronize()
tr = np.zeros(4096, dtype=np.int64)
lib.pdes_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.pdes_tc_trace_read(tr.ctypes.data, 4096))
t0 = tr[2 * 512 + 0]
names = {0: "conv", 1: "mma", 2: "rawi"}
for it in range(7):
    print("tile", it, "epi[wait_start, acc_full, released, done]", [int(tr[3*512 + it*4 + k] - t0) for k in range(4)],
          "mma[acc_empty wait start, end]", [int(tr[3*512+256+it*2+k] - t0) for k in range(2)],
          "conv first chunk start", int(tr[0*512 + it*16*4] - t0))
for it in range(6):
    print("tile", it, "tmem_ld [before, after] x3", [int(tr[3*512+300+it*8+k] - t0) for k in range(6)])
for g in range(14, 34):
    row = []
    for role in (2, 0, 1):
        vals = [int(tr[role * 512 + g * 4 + k] - t0) for k in range(4)]
        row.append(f"{names[role]} {vals}")
    print(g, " | ".join(row))
