"""Where does the tcgen05 K2 spend its time?  Runs the diagnostic build (tools/_bin/libpdes_ablate.so, compiled with
-DPDES_MT_ABLATE) with parts of the kernel switched off: 1 no MMA, 2 no TMA, 4 no convert work, 8 no epilogue stores."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_pde_surrogates_b200 import _native  # noqa: E402

lib = _native.bind(ctypes.CDLL(os.path.join(ROOT, "tools", "_bin", "libpdes_ablate.so")))
dev = torch.device("cuda:0")
Cin, Cout, H, m1, m2 = 193, 192, 96, 10, 10
st = torch.cuda.current_stream().cuda_stream
flush = torch.zeros(64 * 1024 * 1024, device=dev)
for B in (16,):
    w1 = torch.randn(Cin, Cout, m1, m2, dtype=torch.complex64, device=dev)
    w2 = torch.randn(Cin, Cout, m1, m2, dtype=torch.complex64, device=dev)
    wsp = torch.empty(lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2), device=dev)
    _native.check(lib, lib.pdes_mix_tc_pack(w1.data_ptr(), w2.data_ptr(), wsp.data_ptr(), Cin, Cout, H, m1, m2, st))
    X2 = torch.randn(lib.pdes_mix_tc_x2_floats(B, Cin, m1, m2), device=dev)
    O2 = torch.zeros(lib.pdes_mix_tc_o2_floats(B, Cout, m1, m2), device=dev)
    for mask in [int(x) for x in os.environ.get('PDES_MASKS', '0,1,2,4,8,3,5,6,7,15').split(',')]:
        os.environ["PDES_MT_DBG"] = str(mask)
        ts = []
        for it in range(8):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _native.check(lib, lib.pdes_mix_tc_fwd(X2.data_ptr(), wsp.data_ptr(), O2.data_ptr(), B, Cin, Cout, m1, m2, st))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        print(f"B={B:2d} off-mask {mask:2d} ({'MMA ' if mask & 1 else ''}{'TMA ' if mask & 2 else ''}{'convert ' if mask & 4 else ''}{'stores' if mask & 8 else ''}): {ts[len(ts)//2]:8.2f} us", flush=True)

# ---- per-role timeline of CTA 0 (clock64 stamps), full kernel
import numpy as np
os.environ["PDES_MT_DBG"] = os.environ.get("PDES_TRACE_MASK", "0")
_native.check(lib, lib.pdes_mix_tc_fwd(X2.data_ptr(), wsp.data_ptr(), O2.data_ptr(), B, Cin, Cout, m1, m2, st))
torch.cuda.synchronize()
buf = np.zeros(3 * 64 * 4, dtype=np.int64)
lib.pdes_mt_trace_read.argtypes = [ctypes.c_void_p]
lib.pdes_mt_trace_read(buf.ctypes.data)
tr = buf.reshape(3, 64, 4)
t0 = tr[0, 0, 0]
print("chunk | producer: start wait_done issued | convert: start empty_done raw_done arrived | mma: start full_done committed   (cycles since first producer stamp)")
for g in range(36):
    f = lambda a: " ".join(f"{int(x - t0):7d}" if x else "      -" for x in a)
    print(f"{g:3d} | {f(tr[0, g, :3])} | {f(tr[1, g, :4])} | {f(tr[2, g, :3])}")

# ---- does tcgen05 kind::tf32 truncate the low 13 mantissa bits of its operands?  (mask 256 feeds unmasked fp32 as "hi")
os.environ["PDES_MT_DBG"] = "0"
O2.fill_(0)
_native.check(lib, lib.pdes_mix_tc_fwd(X2.data_ptr(), wsp.data_ptr(), O2.data_ptr(), B, Cin, Cout, m1, m2, st))
ref = O2.clone()
os.environ["PDES_MT_DBG"] = "256"
O2.fill_(0)
_native.check(lib, lib.pdes_mix_tc_fwd(X2.data_ptr(), wsp.data_ptr(), O2.data_ptr(), B, Cin, Cout, m1, m2, st))
torch.cuda.synchronize()
print("unmasked-hi run bitwise equal to masked-hi run:", bool(torch.equal(ref, O2)), " max abs diff", float((ref - O2).abs().max()),
      " rel", float((ref - O2).norm() / ref.norm()))
