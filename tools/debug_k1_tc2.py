"""K1 tcgen05 probe: all-ones, row-ramp, column-ramp and random images against numpy, for descriptor variants set by env."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_pde_surrogates_b200 import _native  # noqa: E402

lib = _native.library()
dev = torch.device("cuda:0")
B, C0, H, W, m1, m2 = 1, 4, 96, 64, 10, 10
n = lib.pdes_tables_floats(H, W, m1, m2)
buf = np.zeros(n, dtype=np.float32)
_native.check(lib, lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data))
tab = torch.from_numpy(buf).to(dev)
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(0)
imgs = {"ones": np.ones((H, W)), "h-ramp": np.tile(np.arange(H)[:, None] / H, (1, W)), "w-ramp": np.tile(np.arange(W)[None] / W, (H, 1)),
        "random": rng.standard_normal((H, W))}
variants = [dict(), dict(PDES_K1_DBG="2")] + [dict(PDES_K1_DBG="2", PDES_K1_LBO=str(l), PDES_K1_SBO=str(s_), PDES_K1_LT=str(lt))
                                              for (l, s_, lt) in [(12288, 1024, 1), (512, 12288, 1), (12288, 256, 1)]]
for v in variants:
    for k in ("PDES_K1_DBG", "PDES_K1_LBO", "PDES_K1_SBO", "PDES_K1_LT"):
        os.environ.pop(k, None)
    os.environ.update(v)
    out = []
    for name, img in imgs.items():
        x = torch.zeros(B, C0, H, W, device=dev)
        x[0, 1] = torch.from_numpy(img.astype(np.float32)).to(dev)
        X = torch.full((B, C0, 2 * m1, m2), float("nan"), dtype=torch.complex64, device=dev)
        _native.check(lib, lib.pdes_dft_fwd(x.data_ptr(), C0, None, 0, B, H, W, m1, m2, tab.data_ptr(), 0, X.data_ptr(), st))
        torch.cuda.synchronize()
        F = np.fft.fft2(img.astype(np.float32).astype(np.float64))
        ref = np.concatenate([F[:m1, :m2], F[H - m1:, :m2]], axis=0)
        got = X.cpu().numpy()[0, 1]
        out.append(f"{name} {np.linalg.norm(got - ref) / np.linalg.norm(ref):.3e} (|got| {np.abs(got).mean():.3g})")
    print(v or "default", " | ".join(out), flush=True)
