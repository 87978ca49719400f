import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neural_pde_surrogates_b200 import ops
torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
dev = "cuda:0"
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (B, Cin, N, H, W) in [(16, 193, 192, 96, 64), (16, 192, 192, 94, 62), (16, 385, 192, 100, 68), (8, 193, 192, 96, 64)]:
    conv = torch.nn.Conv2d(Cin, N, 3).to(dev)
    x = torch.randn(B, Cin, H, W, device=dev)
    with torch.no_grad():
        tc = t(lambda: ops.conv3x3_valid(x, conv)) if W % 4 == 0 else float("nan")
        cd = t(lambda: conv(x))
        err = ((ops.conv3x3_valid(x, conv) - conv(x)).norm() / conv(x).norm()).item() if W % 4 == 0 else float("nan")
    fl = 2 * B * N * Cin * 9 * (H - 2) * (W - 2)
    print(f"B={B} {Cin}->{N} {H}x{W}: tcgen05 {tc:8.1f} us ({fl/tc/1e6:6.1f} TFLOP/s fp32-equiv)  cuDNN fp32 {cd:8.1f} us ({fl/cd/1e6:6.1f} TFLOP/s)  rel diff {err:.2e}")
