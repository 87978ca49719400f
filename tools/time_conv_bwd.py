"""cuDNN fp32 (TF32 off, benchmark on) forward / dgrad-only / wgrad-only times of the U-Net's 3x3 valid convs at the
bench shapes, next to the opt-in tcgen05 forward kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = "cuda:0"
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
cb = torch.ops.aten.convolution_backward
for (B, Cin, N, H, W) in [(16, 385, 192, 100, 68), (16, 192, 192, 98, 66), (16, 193, 192, 96, 64), (16, 192, 192, 94, 62),
                          (16, 385, 192, 47, 31), (16, 192, 192, 45, 29), (16, 193, 192, 47, 31)]:
    conv = torch.nn.Conv2d(Cin, N, 3).to(dev)
    x = torch.randn(B, Cin, H, W, device=dev)
    g = torch.randn(B, N, H - 2, W - 2, device=dev)
    w = conv.weight.detach()
    with torch.no_grad():
        fw = t(lambda: conv(x))
        dg = t(lambda: cb(g, x, w, [N], [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False]))
        wg = t(lambda: cb(g, x, w, [N], [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [False, True, True]))
        both = t(lambda: cb(g, x, w, [N], [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, True, True]))
        ops.enable_conv_tc = True
        tc = t(lambda: ops.conv3x3_valid(x, conv)) if W % 4 == 0 else float("nan")
        ops.enable_conv_tc = False
    fl = 2 * B * N * Cin * 9 * (H - 2) * (W - 2) / 1e6
    print(f"B={B} {Cin}->{N} {H}x{W} ({fl/1e3:6.1f} GF): cuDNN fwd {fw:7.1f} us ({fl/fw:5.1f} TF/s) dgrad {dg:7.1f} ({fl/dg:5.1f}) "
          f"wgrad {wg:7.1f} ({fl/wg:5.1f}) both {both:7.1f} | tcgen05 fwd {tc:7.1f} ({fl/tc:5.1f})")
print("dgrad as a forward conv of the padded gradient (ops.ConvValidDgradAsForwardFunction):")
for (B, Cin, N, H, W) in [(16, 385, 192, 100, 68), (16, 192, 192, 98, 66), (16, 193, 192, 96, 64), (16, 385, 192, 47, 31), (16, 192, 192, 45, 29)]:
    conv = torch.nn.Conv2d(Cin, N, 3).to(dev)
    g = torch.randn(B, N, H - 2, W - 2, device=dev)
    w = conv.weight.detach()
    x = torch.randn(B, Cin, H, W, device=dev)
    with torch.no_grad():
        def f():
            wt = w.flip(2, 3).transpose(0, 1).contiguous()
            return torch.nn.functional.conv2d(g, wt, None, padding=2)
        df = t(f)
        ref = cb(g, x, w, [N], [1, 1], [0, 0], [1, 1], False, [0, 0], 1, [True, False, False])[0]
        err = ((f() - ref).norm() / ref.norm()).item()
    fl = 2 * B * N * Cin * 9 * (H - 2) * (W - 2) / 1e6
    print(f"B={B} {Cin}->{N} {H}x{W}: {df:7.1f} us ({fl/df:5.1f} TF/s)  rel diff vs cuDNN dgrad {err:.2e}")
