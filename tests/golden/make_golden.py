"""Generate the golden fixtures by EXECUTING THE UNMODIFIED REFERENCE (build container only).

    python tests/golden/make_golden.py

Imports /root/reference/src (see oracle/reference_loader.py), runs the reference's own SpectralConv2d, FNO_Layer,
UFNO and full activation_wrapper(EncProcDec) model on seeded inputs on the CPU in fp32 and stores inputs, weights,
outputs and gradients as small .npz files next to this script.  The reference ships no tests or golden vectors of
its own, so these files are the pin for the oracle, the torch port and the CUDA path.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.reference_loader import load_reference, twophase_pde  # noqa: E402


def npy(t):
    t = t.detach()
    return t.numpy().copy()


def spectral_cases(ref):
    out = {}
    cases = [(2, 5, 4, 12, 8, 3, 5), (2, 3, 4, 8, 8, 5, 3), (1, 7, 6, 16, 16, 4, 4), (2, 4, 4, 9, 7, 2, 3), (1, 3, 2, 6, 10, 6, 2)]
    for n, (B, Ci, Co, H, W, m1, m2) in enumerate(cases):
        torch.manual_seed(100 + n)
        conv = ref.proc_fno.SpectralConv2d(Ci, Co, (m1, m2))
        with torch.no_grad():                      # O(1) weights so errors are not hidden by the 1/(Ci*Co) init scale
            conv.weights1.mul_(Ci * Co)
            conv.weights2.mul_(Ci * Co)
        x = torch.randn(B, Ci, H, W, requires_grad=True)
        y = conv(x)
        g = torch.randn_like(y)
        (y * g).sum().backward()
        out.update({f"c{n}_x": npy(x), f"c{n}_w1": npy(conv.weights1), f"c{n}_w2": npy(conv.weights2), f"c{n}_y": npy(y),
                    f"c{n}_g": npy(g), f"c{n}_gx": npy(x.grad), f"c{n}_gw1": npy(conv.weights1.grad),
                    f"c{n}_gw2": npy(conv.weights2.grad)})
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "spectral_conv2d.npz"), **out)


def block_case(ref):
    """One UFNO processor (2 blocks) and one FNO processor (1 block), forward + backward."""
    H, W, hf, modes = 16, 12, 8, 3
    pde = twophase_pde(ref, H, W)
    out = {}
    for name, cls in (("ufno", ref.proc_ufno.UFNO), ("fno", ref.proc_fno.FNO)):
        torch.manual_seed(7)
        kw = dict(pde=pde, num_spatial_dims=2, n_cond=1, hidden_features=hf, hidden_blocks=2 if name == "ufno" else 1,
                  fno_modes=modes, padding_mode="circular")
        if name == "ufno":
            kw.update(activation=torch.nn.GELU(), norm=True, ch_mults=[1, 1], is_attn=[False, False], use1x1=True)
        proc = cls(**kw)
        with torch.no_grad():
            for k, p in proc.named_parameters():
                if "weights" in k:
                    p.mul_(hf * hf)                # O(1) spectral weights
        h = torch.randn(2, hf, H, W, requires_grad=True)
        vb = (torch.rand(2, 1, H, W) < 0.3).float()
        y = proc(h=h, variables_broadcast=vb, pos=None)
        g = torch.randn_like(y)
        (y * g).sum().backward()
        out.update({f"{name}_h": npy(h), f"{name}_vb": npy(vb), f"{name}_y": npy(y), f"{name}_g": npy(g), f"{name}_gh": npy(h.grad)})
        for k, v in proc.state_dict().items():
            out[f"{name}_sd_{k}"] = npy(v)
        for k, p in proc.named_parameters():
            out[f"{name}_grad_{k}"] = npy(p.grad)
    np.savez_compressed(os.path.join(HERE, "processors.npz"), **out)


def model_case(ref):
    """Tiny cfg_twophase_ufno-shaped model: one training-style forward/backward and a 5-step rollout through the
    reference trainer's own simulate() (autoregressivepushforwardtrainer.py:288-440)."""
    from neural_pde_surrogates_b200.shell import twophase_model_kwargs
    H, W, B = 24, 16, 2
    pde = twophase_pde(ref, H, W)
    torch.manual_seed(42)
    kw = twophase_model_kwargs("UFNO", hidden_features=16, fno_modes=4, hidden_blocks=2)
    model = ref.models.activation_wrapper(model_class="EncProcDec", **kw, pde=pde)
    u = torch.rand(B, 1, 25, H, W) * 0.5 + 0.1
    mask = (torch.rand(B, 1, H, W) < 0.1).float()
    pos = pde.x[None].repeat(B, 1, 1, 1)
    labels = torch.rand(B, 1, 25, H, W) * 0.5 + 0.1
    args = dict(cond=torch.empty(B, 0), bc=None, pos=pos, t_cond=torch.empty(B, 0), spatial_cond=mask)
    y = model(u, **args)
    loss = torch.sqrt(torch.nn.MSELoss(reduction="sum")(y, labels))     # train_step loss, :157-163
    loss.backward()
    out = {"u": npy(u), "mask": npy(mask), "labels": npy(labels), "y": npy(y), "loss": npy(loss)}
    for k, v in model.state_dict().items():
        out[f"sd_{k}"] = npy(v)
    for k, p in model.named_parameters():
        out[f"grad_{k}"] = npy(p.grad)

    # rollout through the reference trainer's simulate(); a minimal trainer instance without the data plumbing
    T = ref.trainer.AutoregressivePushforwardTrainer
    tr = T.__new__(T)
    from types import SimpleNamespace
    from common.data_creator import DataCreator
    tr.config = SimpleNamespace(device="cpu", process_settings={}, base_resolution=(501, H, W))
    tr.model = model
    tr.data = SimpleNamespace(pde=pde)
    tr.criterion = torch.nn.MSELoss(reduction="sum")
    tr.data_creator = DataCreator(pde=pde, neighbors=3, time_window=25, t_resolution=501, x_resolution=H)
    steps = 5
    u50 = torch.cat([u, torch.zeros_like(u)], dim=2)   # create_data asserts step + tw <= T even for mode="data"
    with torch.no_grad():
        preds = tr.simulate(u50, torch.empty(B, 0), pos, compute_loss=False, include_data=True, nr_gt_steps=1,
                            t_res=25 * (steps + 1), spatial_conditioning=mask, use_bc=False, divide_by_t=False)
    out["rollout"] = np.stack([npy(p) for p in preds[1:]])             # [steps, B, 1, 25, H, W]
    np.savez_compressed(os.path.join(HERE, "model_tiny.npz"), **out)


def trainer_case(ref):
    """The reference trainer's own train_step with a non-zero push-forward unroll (epoch 50 => u in {0,1,2}, drawn from
    the seeded global `random`, autoregressivepushforwardtrainer.py:78-95) + backward, and its test_step
    (:165-286, incl. _test_unrolled_losses :442-514) on whole synthetic trajectories.  Model = model_tiny.npz's."""
    import random
    from types import SimpleNamespace
    from common.data_creator import DataCreator
    from common.interfaces import D
    from neural_pde_surrogates_b200.shell import twophase_model_kwargs
    H, W, B, T = 24, 16, 2, 150
    pde = twophase_pde(ref, H, W)
    torch.manual_seed(42)
    kw = twophase_model_kwargs("UFNO", hidden_features=16, fno_modes=4, hidden_blocks=2)
    model = ref.models.activation_wrapper(model_class="EncProcDec", **kw, pde=pde)      # same init as model_case
    torch.manual_seed(5)
    u_super = torch.rand(B, 1, T, H, W) * 0.5 + 0.1
    mask = (torch.rand(B, 1, H, W) < 0.1).float()
    pos = pde.x[None].repeat(B, 1, 1, 1)
    batch = (torch.empty(0), u_super, pos, torch.empty(B, 0), torch.empty(0), mask)
    Tr = ref.trainer.AutoregressivePushforwardTrainer
    tr = Tr.__new__(Tr)
    tr.config = SimpleNamespace(device="cpu", batch_size=B, lr_step_interval=25, unrolling=8, process_settings={},
                                base_resolution=(T, H, W), time_window=25, nr_gt_steps=1)
    tr.model, tr.criterion = model, torch.nn.MSELoss(reduction="sum")
    tr.data = SimpleNamespace(pde=pde, data_interface=D.sim2d)
    tr.data_creator = DataCreator(pde=pde, neighbors=3, time_window=25, t_resolution=T, x_resolution=H)
    out = {"u_super": npy(u_super), "mask": npy(mask), "epoch": np.array(50)}
    for seed in (3, 7, 5):                                    # seeds chosen to cover unroll counts 0, 1 and 2
        random.seed(seed)
        st = random.getstate()
        unrolled = random.choice(list(range(3)))
        random.setstate(st)
        model.zero_grad()
        loss, pred = tr.train_step(batch, 50, 0, None)
        loss.backward()
        out[f"ts{seed}_unrolled"] = np.array(unrolled)
        out[f"ts{seed}_loss"] = npy(loss)
        out[f"ts{seed}_pred"] = npy(pred)
        for k, p in model.named_parameters():
            if k.endswith("conv.weights1") or k.endswith("w.weight") or k.startswith("encoder.encoder.0.weight"):
                out[f"ts{seed}_grad_{k}"] = npy(p.grad)
    with torch.no_grad():
        val, info = tr.test_step(batch, 0)
    out["test_loss"] = npy(val)
    for k, v in info.items():
        out["test_info_" + k.replace(" ", "_").replace(",", "")] = npy(torch.as_tensor(v))
    np.savez_compressed(os.path.join(HERE, "trainer_steps.npz"), **out)
    print("trainer_case unroll counts:", {s: int(out[f"ts{s}_unrolled"]) for s in (3, 7, 5)})


if __name__ == "__main__":
    torch.set_num_threads(4)
    ref = load_reference()
    spectral_cases(ref)
    block_case(ref)
    model_case(ref)
    trainer_case(ref)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
