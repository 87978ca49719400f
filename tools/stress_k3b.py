"""Bitwise run-to-run determinism of K3a (old + new) and K3b at the bench shape: exposes timing-dependent bugs and
prints where the differing outputs sit (GPU box only)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from backends import CudaBackend  # noqa: E402

be = CudaBackend()
lib, p, ck = be.lib, be.ptr, be.check
n_it = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B, C0, C1, Cout, H, W, m1, m2 = 16, 192, 1, 192, 96, 64, 10, 10
Cin = C0 + C1
rng = np.random.default_rng(3)
dev = be.dev
P = torch.randn(2, B, Cout, 2 * m1, m2, dtype=torch.complex64, device=dev)
x0 = torch.randn(B, C0, H, W, device=dev)
x1 = torch.randn(B, C1, H, W, device=dev)
wc = torch.randn(Cout, Cin, device=dev) / Cin ** 0.5
bias = torch.randn(Cout, device=dev)
res = torch.randn(B, Cout, H, W, device=dev)
tab = be.tables(H, W, m1, m2)
st = be.stream
pack = be.empty((lib.pdes_gemm_tc_pack_floats(Cin, Cout),))
ck(lib.pdes_gemm_tc_pack_t(p(wc), Cin, Cin, Cout, p(pack), st))
flush = torch.zeros(64 * 1024 * 1024, device=dev)


def run_k3a():
    Z = be.empty((B, H, 2 * m2, Cout))
    ck(lib.pdes_inv_h(p(P), 2, B, Cout, H, m1, m2, p(tab), p(Z), st))
    return Z


def run_k3b(Z):
    out = be.empty((B, Cout, H, W))
    pre = be.empty((B, Cout, H, W))
    ck(lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(x0), C0, p(x1), C1, p(bias), p(res), p(tab), 0, p(out), p(pre),
                              B, Cout, H, W, m1, m2, 1, st))
    return out, pre


def where(d, shape_names):
    idx = torch.nonzero(d)
    print("   differing elements:", idx.shape[0])
    for k, nm in enumerate(shape_names):
        u = torch.unique(idx[:, k])
        print("   ", nm, u[:12].tolist(), "... n =", u.numel())


Z0 = run_k3a()
out0, pre0 = run_k3b(Z0)
torch.cuda.synchronize()
bad_a = bad_b = 0
for i in range(n_it):
    if i % 2:
        flush.add_(1.0)
    Z = run_k3a()
    if not torch.equal(Z.view(torch.int32), Z0.view(torch.int32)):
        bad_a += 1
        print("K3a differs at iteration", i)
        where(Z.view(torch.int32) != Z0.view(torch.int32), ["b", "h", "j", "n"])
    out, pre = run_k3b(Z0)
    if not torch.equal(pre.view(torch.int32), pre0.view(torch.int32)):
        bad_b += 1
        print("K3b differs at iteration", i)
        d = pre.view(torch.int32) != pre0.view(torch.int32)
        where(d, ["b", "n", "h", "w"])
        dd = (pre - pre0)[d]
        print("    |diff| max", float(dd.abs().max()), "mean", float(dd.abs().mean()), "nan", int(torch.isnan(dd).sum()))
        idx = torch.nonzero(d)
        b_, h_ = int(idx[0, 0]), int(idx[0, 2])
        w0 = int(idx[:, 3].min()) // 32 * 32
        D = (pre - pre0)[b_, :, h_, w0:w0 + 32].double()              # [n][32 px]
        U, S, Vh = torch.linalg.svd(D)
        print("    singular values", [round(float(x), 4) for x in S[:4]])
        u = U[:, 0]
        xin = torch.cat([x0, x1], 1)
        corr = (wc.double().t() @ u) / wc.double().norm(dim=0)       # |cos| between u and column k of the 1x1 weight
        k = int(corr.abs().argmax())
        print("    best k", k, "cos", round(float(corr[k]), 4), " tile t =", b_ * 48 + h_ // 2, "cta", (b_ * 48 + h_ // 2) % 148,
              "it", (b_ * 48 + h_ // 2) // 148)
        delta = (D.t() @ wc.double()[:, k]) / (wc.double()[:, k] ** 2).sum()    # per-pixel error of A[:, k]
        xt = xin[b_, k, h_, w0:w0 + 32].double()
        print("    delta[:6]", [round(float(v), 4) for v in delta[:6]], " x_true[:6]", [round(float(v), 4) for v in xt[:6]])
        stale = delta + xt
        t_ = b_ * 48 + h_ // 2
        for name, tt, kk in (("same tile, k+128 (later user of the slot)", t_, k + 128), ("previous tile, k+128 (previous content)", t_ - 148, k + 128),
                             ("previous tile, same k", t_ - 148, k), ("next tile, same k", t_ + 148, k), ("same tile k-128", t_, k - 128)):
            if 0 <= tt < B * 48 and 0 <= kk < Cin:
                cand = xin[tt // 48, kk, (tt % 48) * 2 + (h_ & 1), w0:w0 + 32].double()
                print("     ", name, "max |stale - cand| =", round(float((stale - cand).abs().max()), 5))
        # where does the stale value come from?  search the same channel for a matching run of values
        flat = xin[:, k].reshape(-1).double()
        m = torch.nonzero((flat - stale[0]).abs() < 1e-4).flatten()
        for j in m[:4].tolist():
            ok = j + 32 <= flat.numel() and bool(((flat[j:j + 32] - stale).abs() < 1e-3).all())
            print("    stale run found at b,h,w =", j // (H * W), (j % (H * W)) // W, j % W, "full match" if ok else "first only")
print("done", n_it, "K3a bad", bad_a, "K3b bad", bad_b)
