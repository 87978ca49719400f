#!/usr/bin/env python
"""Benchmark of the B200-native U-FNO hot path (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

Workload (`config.workload`): cfg_twophase_ufno -- U-FNO, 3 blocks, width 192, modes 10x10, grid 96x64, tw 25,
Cin = 193 (mask-only conditioning), per-GPU batch 16, fp32 (cuDNN TF32 off), push-forward unroll u=0, Adam.
A step = one optimizer step (forward + backward + gradient all-reduce + Adam) on one batch of synthetic windows.
`value`   : training samples/s, whole job, inputs already resident in HBM.
`e2e`     : the same through the public trainer API with the batch in pinned HOST memory (H2D of the windows and
            D2H of the loss inside the timed region).
`roofline`: the fused spectral-block forward chain (K1+K2+K3a+K3b), algorithmic bytes / CUDA-event time measured
            in situ during the timed steps, against the measured HBM copy bandwidth.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

H, W, TW, WIDTH, MODES, BLOCKS, NCOND = 96, 64, 25, 192, 10, 3, 1


def block_bytes(B, Cin=WIDTH + NCOND, Cout=WIDTH):
    """Algorithmic bytes of one fused U-FNO block tail, fp32 (SURVEY.md §8d / BASELINE.md §3)."""
    spec = 4 * B * Cin * H * W + 16 * Cin * Cout * MODES * MODES + 4 * B * Cout * H * W
    return spec + 4 * B * Cout * H * W + 4 * Cout * Cin + 4 * Cout


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def synthetic_batch(B, pde, device, gen):
    """SURVEY.md §8d synthetic inputs: u ~ U(0.1, 0.6), Bernoulli(0.1) obstacle mask, empty static conditioning."""
    u = torch.rand(B, 1, TW, H, W, generator=gen) * 0.5 + 0.1
    labels = torch.rand(B, 1, TW, H, W, generator=gen) * 0.5 + 0.1
    mask = (torch.rand(B, 1, H, W, generator=gen) < 0.1).float()
    pos = pde.x[None].repeat(B, 1, 1, 1)
    return u, labels, mask, pos


def build(device, seed=42, hidden_blocks=BLOCKS):
    import neural_pde_surrogates_b200 as npb
    torch.manual_seed(seed)
    pde = npb.TwoPhasePDE(H, W)
    model = npb.build_twophase_model(pde=pde, hidden_features=WIDTH, fno_modes=MODES, hidden_blocks=hidden_blocks)
    return model.to(device), pde


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own algorithm on the host cores: the CPU port of the same model (oracle/torch_port.py; the
    unmodified reference cannot travel to the GPU box).  Bounded sample of the same workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer
    from oracle.torch_port import cpu_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, pde = build("cpu")
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    gen = torch.Generator().manual_seed(1)
    Bs = 4
    with cpu_port():
        while True:
            tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device="cpu", batch_size=Bs)
            u, labels, mask, pos = synthetic_batch(Bs, pde, "cpu", gen)
            t0 = time.perf_counter()
            loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(Bs, 0), mask)
            tr.optimizer_step(loss)
            t_probe = time.perf_counter() - t0
            if Bs == 1 or t_probe * (args.steps + args.warmup) < 200:
                break
            Bs //= 2
        for _ in range(max(args.warmup - 1, 0)):
            loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(Bs, 0), mask)
            tr.optimizer_step(loss)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(Bs, 0), mask)
            tr.optimizer_step(loss)
            float(loss.detach())
        dt = time.perf_counter() - t0
    v = Bs * args.steps / dt
    sample = f"{args.steps} optimizer steps of cfg_twophase_ufno at batch {Bs} (fwd+bwd+Adam, u=0) on the host CPU"
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(Bs, 1),
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(B, n):
    return {"workload": "cfg_twophase_ufno train step (U-FNO x3, width 192, modes 10x10, grid 96x64, tw 25, Cin 193, Adam, unroll u=0)",
            "per_gpu_batch": B, "global_batch": B * n, "parallelism": f"dp{n}", "grid": [H, W], "precision": "fp32 (cuDNN TF32 off, cudnn.benchmark on; spectral block 3xTF32 split on tcgen05 = fp32-faithful)",
            "l2_policy": "inputs+weights+activations per step (~1 GB) exceed the 126 MB L2; no explicit flush"}


def cpu_baseline_sample():
    """Oracle port timed on the host cores on a bounded sample (one warm-up + timed steps at batch 4)."""
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer
    from oracle.torch_port import cpu_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, pde = build("cpu")
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    gen = torch.Generator().manual_seed(1)
    Bs, n = 4, 0
    tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device="cpu", batch_size=Bs)
    u, labels, mask, pos = synthetic_batch(Bs, pde, "cpu", gen)
    with cpu_port():
        loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(Bs, 0), mask)
        tr.optimizer_step(loss)
        t0 = time.perf_counter()
        while n < 2 or (time.perf_counter() - t0 < 12 and n < 8):
            loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(Bs, 0), mask)
            tr.optimizer_step(loss)
            n += 1
        dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} optimizer steps of the same model at batch {Bs} with the CPU port of the reference algorithm "
                      f"(oracle/torch_port.py: torch.fft + einsum + conv), {dt:.1f} s"}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    from neural_pde_surrogates_b200 import dp, ops
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer

    rank, world, local = dp.init_distributed()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.allow_tf32 = args.tf32_convs
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = not args.no_cudnn_benchmark
    B = args.batch

    model, pde = build(dev)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)                 # defaults/optimizer.py:4-7
    tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device=dev, batch_size=B)
    if world > 1:
        dp.make_data_parallel(tr, seed=42)
    gen = torch.Generator().manual_seed(1234 + rank)
    u, labels, mask, pos = synthetic_batch(B, pde, dev, gen)
    u_h, labels_h = u.pin_memory(), labels.pin_memory()                  # host copies for the e2e leg
    u_d, labels_d, mask_d, pos_d = u.to(dev), labels.to(dev), mask.to(dev), pos.to(dev)
    cond = torch.empty(B, 0, device=dev)

    def step_resident():
        loss, _ = tr.train_step_windows(u_d, labels_d, pos_d, cond, mask_d)
        tr.optimizer_step(loss)
        return loss

    def step_e2e():
        ud = u_h.to(dev, non_blocking=True)
        ld = labels_h.to(dev, non_blocking=True)
        loss, _ = tr.train_step_windows(ud, ld, pos_d, cond, mask_d)
        tr.optimizer_step(loss)
        return float(loss.detach())                                               # D2H read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step_resident()
    ops.reset_counters()
    ops.enable_timing(True)
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    ops.enable_timing(False)
    launches = ops.counters()["launches"]
    chain = ops.collect_timings()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    value = B * world * args.steps / (ms * 1e-3)
    e2e_value = B * world * args.steps / (ms_e2e * 1e-3)

    extra = {}
    if rank == 0 and not args.no_extras:
        extra = extras(tr, model, pde, dev, B, args)
    if rank == 0:
        peak, peak_src = measured_peak()
        fwd_us = statistics.mean(chain["block_forward"]) * 1e3 if chain["block_forward"] else float("nan")
        bwd_us = statistics.mean(chain["block_backward"]) * 1e3 if chain["block_backward"] else float("nan")
        ach = block_bytes(B) / (fwd_us * 1e-6) / 1e9
        line = {"metric": "train_samples_per_s", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(B, world),
                "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": int(u_h.numel() * 4 + labels_h.numel() * 4),
                        "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches,
                "clocks": clocks.summary(),
                "roofline": {"bound": "hbm", "kernel": "fno_block_forward chain (K1 k_dft_fwd_fast + K2 k_mix_tma + K3a k_inv_h + weight pack + K3b k_inv_w_gemm_tc_v3 on tcgen05)",
                             "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                             # dram__bytes_read+write summed over the chain's kernels, one ncu --set full capture at B=16
                             # (profiles/r01_v5_ncu_full_block_B16.txt); above the algorithmic bytes because h is read by
                             # K1 and again by K3b (+76 MB), the training forward also stores the pre-activation (+75 MB)
                             # and the Z / partial-sum intermediates are not fully L2-resident
                             "traffic": 432.0e6 if B == 16 else None,
                             "peak_source": peak_src, "alg_bytes_per_launch": block_bytes(B), "us_per_launch": fwd_us,
                             "launches_timed": len(chain["block_forward"]), "block_backward_us": bwd_us},
                "cpu_baseline": extra.pop("cpu_baseline", None)}
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(tr, model, pde, dev, B, args):
    """Secondary numbers of the same run: rollout steps/s (eager and CUDA-graph), block micro-benchmark, CPU baseline."""
    out = {}
    model.eval()
    Br = args.rollout_batch
    gen = torch.Generator().manual_seed(7)
    u, _, mask, pos = synthetic_batch(Br, pde, dev, gen)
    u, pos = u.to(dev), pos.to(dev)
    mask = torch.zeros_like(mask).to(dev)                                 # twophase_no_obstacle: mask all zeros
    cond = torch.empty(Br, 0, device=dev)
    nsteps = 50
    kw = dict(compute_loss=False, include_data=True, nr_gt_steps=1, t_res=TW * (nsteps + 1), spatial_conditioning=mask,
              use_bc=False, divide_by_t=False)
    res = {}
    with torch.no_grad():
        for graph in (False, True):
            tr.simulate(u, cond, pos, graph=graph, **dict(kw, t_res=TW * 3))      # warm-up (and graph capture)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            preds = tr.simulate(u, cond, pos, graph=graph, **kw)
            e1.record()
            torch.cuda.synchronize()
            res["graph" if graph else "eager"] = Br * nsteps / (e0.elapsed_time(e1) * 1e-3)
        finite = bool(torch.isfinite(preds[-1]).all())
    out["rollout"] = {"metric": "rollout_trajectory_steps_per_s", "steps": nsteps, "trajectories": Br,
                      "eager": res["eager"], "cuda_graph": res["graph"], "finite": finite,
                      "note": "one step = one model application advancing 25 frames; state stays in HBM"}
    model.train()
    # push-forward training step with the maximum unroll of the shipped config (u = 8 no-grad applications feeding one
    # differentiable application, autoregressivepushforwardtrainer.py:115-144); the headline `value` is u = 0
    ut, lt, mt, pt = synthetic_batch(B, pde, dev, torch.Generator().manual_seed(11))
    ut, lt, mt, pt = ut.to(dev), lt.to(dev), mt.to(dev), pt.to(dev)
    condt = torch.empty(B, 0, device=dev)

    def step_u8():
        loss, _ = tr.train_step_windows(ut, lt, pt, condt, mt, unrolled=8, next_labels=lambda k: lt)
        tr.optimizer_step(loss)
    try:                                                         # a secondary number must never cost the headline line
        for _ in range(2):
            step_u8()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n8 = 3
        e0.record()
        for _ in range(n8):
            step_u8()
        e1.record()
        torch.cuda.synchronize()
        ms8 = e0.elapsed_time(e1) / n8
        out["train_unroll8"] = {"metric": "train_samples_per_s", "unroll": 8, "value": B / (ms8 * 1e-3), "ms_per_step": ms8,
                                "note": "per GPU; 8 no-grad model applications + 1 with grad per optimizer step"}
    except Exception as exc:                                     # noqa: BLE001
        out["train_unroll8"] = {"error": repr(exc)[:200]}
    if not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline_sample()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="per-GPU batch (config default, defaults/base.py:6)")
    ap.add_argument("--rollout-batch", type=int, default=16)
    ap.add_argument("--tf32-convs", action="store_true", help="let cuDNN use TF32 in the U-Net branch (reported separately)")
    ap.add_argument("--no-cudnn-benchmark", action="store_true",
                    help="do not let cuDNN autotune the conv algorithms of the U-Net branch (default: autotune on)")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device for the b200 arm (there is no CPU fallback); "
                             "use --impl reference for the host baseline")
        run_ours(args)


if __name__ == "__main__":
    main()
