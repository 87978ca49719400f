"""MMA-issue-thread trace of the persistent K3b kernel (library built with PDES_NVCC_EXTRA=-DPDES_TC_TRACE=2)."""
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
ACT = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(3):
    lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, ACT, st)
torch.cuda.synchronize()
tr = np.zeros(4096, dtype=np.int64)
lib.pdes_tc_trace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
print("rc", lib.pdes_tc_trace_read(tr.ctypes.data, 4096))
t0 = tr[0]
print("chunk: start | +a_full +b_full | 6 MMA issue deltas | +commits || period")
prev = None
rows = []
for g in range(16, 80):
    v = tr[g * 10:(g + 1) * 10] - t0
    d = np.diff(v)
    print(f"{g:3d}: {int(v[0]):7d} | {int(d[0]):4d} {int(d[1]):4d} | " + " ".join(f"{int(x):4d}" for x in d[2:8]) + f" | {int(d[8]):4d} || {int(v[0] - prev) if prev is not None else 0}")
    rows.append(int(v[0] - prev) if prev is not None else 0)
    prev = v[0]
print('mean period', np.mean(rows[1:]), 'tile periods', [int(tr[(16*k)*10] - tr[(16*(k-1))*10]) for k in range(2, 6)])
