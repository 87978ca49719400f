"""ncu driver: the tcgen05 K3b kernel alone (forward epilogue variant) at the config shape."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import _native
lib = _native.library()
if os.environ.get("PDES_MODE"):
    lib.pdes_set_tensor_core_mode(int(os.environ["PDES_MODE"]))
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
variant = sys.argv[2] if len(sys.argv) > 2 else "full"
C0, C1, Cout, H, W, m1, m2 = 192, 1, 192, 96, 64, 10, 10
Cin = C0 + C1
n = lib.pdes_tables_floats(H, W, m1, m2)
buf = np.zeros(n, dtype=np.float32); lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data)
tab = torch.from_numpy(buf).to(dev)
h = torch.randn(B, C0, H, W, device=dev); vb = torch.randn(B, C1, H, W, device=dev)
res = torch.randn(B, Cout, H, W, device=dev); bias = torch.randn(Cout, device=dev)
wct = torch.randn(Cin, Cout, device=dev) / Cin ** 0.5
Z = torch.randn(B, H, 2 * m2, Cout, device=dev)
out = torch.empty(B, Cout, H, W, device=dev)
pack = torch.empty(lib.pdes_gemm_tc_pack_floats(Cin, Cout), device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: None if t is None else t.data_ptr()
_native.check(lib, lib.pdes_gemm_tc_pack(p(wct), Cout, Cin, Cout, p(pack), st))
for _ in range(3):
    if variant == "full":
        _native.check(lib, lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 1, st))
    elif variant == "nospec":
        _native.check(lib, lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 1, st))
    else:  # bare GEMM
        _native.check(lib, lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, p(vb), C1, None, None, p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 0, st))
torch.cuda.synchronize()
for v in ("full", "nospec", "bare", "nostore"):
    fn = {"full": lambda: lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 1, st),
          "nospec": lambda: lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 1, st),
          "nostore": lambda: lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, p(vb), C1, None, None, p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 77, st),
          "bare": lambda: lib.pdes_inv_w_gemm_tc(None, p(pack), p(h), C0, p(vb), C1, None, None, p(tab), 0, p(out), None, B, Cout, H, W, m1, m2, 0, st)}[v]
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    print(v, e0.elapsed_time(e1) / 5 * 1e3, "us")
