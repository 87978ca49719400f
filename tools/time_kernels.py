"""Time every kernel of the spectral block with CUDA events at the config shapes (GPU box only)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neural_pde_surrogates_b200 import _native  # noqa: E402

lib = _native.library()
dev = torch.device("cuda:0")
PEAK = 6545.6
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def tables(H, W, m1, m2):
    n = lib.pdes_tables_floats(H, W, m1, m2)
    buf = np.zeros(n, dtype=np.float32)
    _native.check(lib, lib.pdes_tables_fill(H, W, m1, m2, buf.ctypes.data))
    return torch.from_numpy(buf).to(dev)


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    out = {}
    flush = torch.zeros(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > L2
    for B in (4, 16):
        C0, C1, Cout, H, W, m1, m2 = 192, 1, 192, 96, 64, 10, 10
        Cin, HW, MM = C0 + C1, H * W, m1 * m2
        st = torch.cuda.current_stream().cuda_stream
        tab = tables(H, W, m1, m2)
        h = torch.randn(B, C0, H, W, device=dev)
        vb = torch.randn(B, C1, H, W, device=dev)
        res = torch.randn(B, Cout, H, W, device=dev)
        g = torch.randn(B, Cout, H, W, device=dev)
        w1 = torch.randn(Cin, Cout, m1, m2, dtype=torch.complex64, device=dev) / Cin
        w2 = torch.randn(Cin, Cout, m1, m2, dtype=torch.complex64, device=dev) / Cin
        wc = torch.randn(Cout, Cin, device=dev) / Cin ** 0.5
        wct = wc.t().contiguous()
        bias = torch.randn(Cout, device=dev)
        X = torch.empty(B, Cin, 2 * m1, m2, dtype=torch.complex64, device=dev)
        ns = lib.pdes_mix_suggest_splits(B, Cin, Cout, m1, m2)
        P = torch.empty(ns, B, Cout, 2 * m1, m2, dtype=torch.complex64, device=dev)
        Z = torch.empty(B, H, 2 * m2, Cout, device=dev)
        outp = torch.empty(B, Cout, H, W, device=dev)
        pre = torch.empty(B, Cout, H, W, device=dev)
        ws = torch.empty(lib.pdes_block_fwd_workspace_floats(B, Cin, Cout, H, W, m1, m2), device=dev)
        wsb = torch.empty(lib.pdes_block_bwd_workspace_floats(B, C0, C1, Cout, H, W, m1, m2), device=dev)
        pack = torch.empty(lib.pdes_gemm_tc_pack_floats(Cin, Cout), device=dev)
        gpre = torch.empty_like(g)
        dh = torch.empty_like(h)
        gw1, gw2 = torch.empty_like(w1), torch.empty_like(w2)
        dwc, dbias = torch.empty_like(wc), torch.empty_like(bias)
        GO = torch.empty(B, Cout, 2 * m1, m2, dtype=torch.complex64, device=dev)
        wgtc = torch.empty(lib.pdes_wgrad_tc_workspace_floats(Cout, Cin), device=dev)
        wgws = torch.empty(lib.pdes_wgrad_workspace_floats(B, Cout, Cin, HW), device=dev)
        p = lambda t: t.data_ptr()
        ck = lambda c: _native.check(lib, c)
        wsp = torch.empty(lib.pdes_mix_tc_pack_floats(Cin, Cout, m1, m2), device=dev)
        X2 = torch.zeros(lib.pdes_mix_tc_x2_floats(B, Cin, m1, m2), device=dev)
        O2 = torch.zeros(lib.pdes_mix_tc_o2_floats(B, Cout, m1, m2), device=dev)
        ck(lib.pdes_mix_tc_pack(p(w1), p(w2), p(wsp), Cin, Cout, H, m1, m2, st))
        pack1 = torch.empty(lib.pdes_gemm_tc_pack_floats(Cin, Cout), device=dev)
        ck(lib.pdes_gemm_tc_pack_t(p(wc), Cin, Cin, Cout, p(pack1), st))
        tc_runs = {
            "K1_dft_fwd2_modemajor": (lambda: ck(lib.pdes_dft_fwd2(p(h), C0, p(vb), C1, B, H, W, m1, m2, p(tab), 0, p(X), p(X2), st)),
                                      4 * B * Cin * HW + 16 * B * Cin * 2 * MM),
            "K2_mix_tcgen05": (lambda: ck(lib.pdes_mix_tc_fwd(p(X2), p(wsp), p(O2), B, Cin, Cout, m1, m2, st)),
                               16 * Cin * Cout * MM + 8 * B * Cin * 2 * MM + 8 * B * Cout * 2 * MM),
            "K3a_inv_h_modes": (lambda: ck(lib.pdes_inv_h_modes(p(O2), B, Cout, H, m1, m2, p(tab), p(Z), st)),
                                8 * B * Cout * 2 * MM + 4 * B * H * 2 * m2 * Cout),
            "K2_dx_tcgen05": (lambda: ck(lib.pdes_mix_tc_dx(p(X2), p(wsp), p(O2), B, Cin, Cout, C0, m1, m2, st)),
                              16 * Cin * Cout * MM + 8 * B * Cout * 2 * MM + 8 * B * C0 * 2 * MM),
            "spectral_weight_pack": (lambda: ck(lib.pdes_mix_tc_pack(p(w1), p(w2), p(wsp), Cin, Cout, H, m1, m2, st)),
                                     32 * Cin * Cout * MM),
            "block_forward_tc_cached_packs": (lambda: ck(lib.pdes_block_forward(p(h), C0, p(vb), C1, p(w1), p(w2), p(wsp), p(wc), p(pack1), p(bias),
                                                                                p(res), p(tab), p(X), p(ws), p(outp), None, B, Cout, H, W, m1, m2, 1, st)),
                                              4 * B * Cin * HW + 16 * Cin * Cout * MM + 8 * B * Cout * HW + 4 * Cout * Cin + 4 * Cout),
            "block_forward_tc_with_pre": (lambda: ck(lib.pdes_block_forward(p(h), C0, p(vb), C1, p(w1), p(w2), p(wsp), p(wc), p(pack1), p(bias),
                                                                            p(res), p(tab), p(X), p(ws), p(outp), p(pre), B, Cout, H, W, m1, m2, 1, st)),
                                          4 * B * Cin * HW + 16 * Cin * Cout * MM + 8 * B * Cout * HW + 4 * Cout * Cin + 4 * Cout),
        }
        runs = {
            "K1_dft_fwd": (lambda: ck(lib.pdes_dft_fwd(p(h), C0, p(vb), C1, B, H, W, m1, m2, p(tab), 0, p(X), st)),
                           4 * B * Cin * HW + 8 * B * Cin * 2 * MM),
            "K2_mix_fwd": (lambda: ck(lib.pdes_mix_fwd(p(X), p(w1), p(w2), p(P), ns, B, Cin, Cout, H, m1, m2, st)),
                           16 * Cin * Cout * MM + 8 * B * Cin * 2 * MM + 8 * ns * B * Cout * 2 * MM),
            "K3a_inv_h": (lambda: ck(lib.pdes_inv_h(p(P), ns, B, Cout, H, m1, m2, p(tab), p(Z), st)),
                          8 * ns * B * Cout * 2 * MM + 4 * B * H * 2 * m2 * Cout),
            "K3b_inv_w_gemm": (lambda: ck(lib.pdes_inv_w_gemm(p(Z), p(wct), Cout, p(h), C0, p(vb), C1, p(bias), p(res),
                                                             p(tab), 0, p(outp), None, B, Cout, H, W, m1, m2, 1, st)),
                               4 * B * Cin * HW + 8 * B * Cout * HW + 4 * B * H * 2 * m2 * Cout),
            "K3b_tcgen05_3xtf32": (lambda: (ck(lib.pdes_gemm_tc_pack(p(wct), Cout, Cin, Cout, p(pack), st)),
                                            ck(lib.pdes_inv_w_gemm_tc(p(Z), p(pack), p(h), C0, p(vb), C1, p(bias), p(res), p(tab), 0,
                                                                      p(outp), None, B, Cout, H, W, m1, m2, 1, st))),
                                   4 * B * Cin * HW + 8 * B * Cout * HW + 4 * B * H * 2 * m2 * Cout),
            "block_forward": (lambda: ck(lib.pdes_block_forward(p(h), C0, p(vb), C1, p(w1), p(w2), None, p(wc), None, p(bias), p(res),
                                                                p(tab), p(X), p(ws), p(outp), None, B, Cout, H, W, m1, m2, 1, st)),
                              4 * B * Cin * HW + 16 * Cin * Cout * MM + 8 * B * Cout * HW + 4 * Cout * Cin + 4 * Cout),
            "act_bwd": (lambda: ck(lib.pdes_act_bwd(p(g), p(pre), p(gpre), g.numel(), 1, st)), 12 * B * Cout * HW),
            "mix_dw": (lambda: ck(lib.pdes_mix_dw(p(X), p(GO), p(gw1), p(gw2), B, Cin, Cout, H, m1, m2, st)),
                       16 * Cin * Cout * MM),
            "wgrad": (lambda: ck(lib.pdes_wgrad(p(g), p(h), C0, p(vb), C1, p(dwc), p(dbias), p(wgws), B, Cout, HW, st)),
                      4 * B * (Cin + Cout) * HW),
            "wgrad_tcgen05": (lambda: ck(lib.pdes_wgrad_tc(p(g), p(h), C0, p(vb), C1, p(dwc), p(dbias), p(wgtc), B, Cout, HW, st)),
                              4 * B * (Cin + Cout) * HW),
            "block_backward": (lambda: ck(lib.pdes_block_backward(p(g), p(pre), p(h), C0, p(vb), C1, p(X), p(w1), p(w2), p(wsp), p(wc), None,
                                                                  p(tab), p(wsb), p(gpre), p(dh), p(gw1), p(gw2), p(dwc), p(dbias),
                                                                  B, Cout, H, W, m1, m2, 1, st)),
                               4 * B * (3 * Cout + 2 * Cin + C0) * HW + 32 * Cin * Cout * MM),
        }
        pre.copy_(torch.randn_like(pre))
        only = os.environ.get("PDES_TIME_ONLY")
        if lib.pdes_mix_tc_ok(B, Cin, Cout, m1, m2):
            runs = {**tc_runs, **runs}
        for name, (fn, nbytes) in runs.items():
            if only and not any(t in name for t in only.split(",")):
                continue
            us = timeit(fn, flush=flush)
            us_hot = timeit(fn)
            gbs = nbytes / us / 1e3
            out[f"B{B}.{name}"] = dict(us_cold=round(us, 2), us_hot=round(us_hot, 2), alg_MB=round(nbytes / 1e6, 2),
                                       GBps=round(gbs, 1), frac_of_measured_hbm=round(gbs / PEAK, 3))
            print(f"B={B:2d} {name:16s} cold {us:9.2f} us  hot {us_hot:9.2f} us  {nbytes/1e6:8.2f} MB  {gbs:8.1f} GB/s  {gbs/PEAK:6.3f}")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("PDES_TIME_OUT", "kernel_times.json")), "w"), indent=1)


if __name__ == "__main__":
    main()
