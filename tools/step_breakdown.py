"""Kernel-time breakdown of one bench.py training step (cudnn.benchmark on) from torch.profiler (CUPTI, no replay).

usage: python tools/step_breakdown.py [batch] > profiles/<name>.txt
"""
import os
import sys
from collections import defaultdict

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
model, pde = bench.build(dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device=dev, batch_size=B)
gen = torch.Generator().manual_seed(1234)
u, labels, mask, pos = bench.synthetic_batch(B, pde, dev, gen)
u, labels, mask, pos = u.to(dev), labels.to(dev), mask.to(dev), pos.to(dev)
cond = torch.empty(B, 0, device=dev)


def step():
    loss, _ = tr.train_step_windows(u, labels, pos, cond, mask)
    tr.optimizer_step(loss)


for _ in range(4):
    step()
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step()
    torch.cuda.synchronize()

tot = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        t = tot[ev.name[:110]]
        t[0] += 1
        t[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
total = sum(v[1] for v in tot.values())
print(f"one training step, B={B}, cudnn.benchmark on: {sum(v[0] for v in tot.values())} launches, {total / 1e3:.2f} ms device time")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{name:112s} {n:5d} {us / 1e3:9.3f} ms {100 * us / total:5.1f}%")
ours = sum(v[1] for k, v in tot.items() if "pdes::" in k)
print(f"\nour kernels (pdes::*): {ours / 1e3:.3f} ms = {100 * ours / total:.2f}% of the step")

# per-op (CPU-side aten op) attribution of device time for the convolution calls, by input shape
print("\nconvolution calls by shape (device time of the aten op, forward and backward):")
rows = defaultdict(lambda: [0, 0.0])
for ev in prof.key_averages(group_by_input_shape=True):
    if "conv" in ev.key and ("cudnn" in ev.key or "convolution_backward" in ev.key or ev.key == "aten::convolution"):
        dt = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        rows[(ev.key, str(ev.input_shapes)[:150])] = [ev.count, dt]
for (k, shp), (n, us) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{k:40s} {n:4d} {us / 1e3:9.3f} ms  {shp}")
