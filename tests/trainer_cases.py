"""Shared bodies of the trainer parity tests (CPU with the torch port, GPU with the CUDA path) against
tests/golden/trainer_steps.npz = outputs of the UNMODIFIED reference trainer's own train_step / test_step
(tests/golden/make_golden.py:trainer_case)."""
import random

import torch

from parity_util import golden, rel_l2, tiny_model
from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer

SEEDS = (3, 7, 5)          # push-forward unroll counts 0, 1, 2 at epoch 50


def _setup(device):
    model, pde, _ = tiny_model(device)
    g = golden("trainer_steps.npz")
    B, _, T, H, W = g["u_super"].shape
    u_super = torch.from_numpy(g["u_super"])                          # whole trajectories stay on the host (base.py:487-489 moves them)
    mask = torch.from_numpy(g["mask"]).to(device)
    pos = pde.x.to(device)[None].repeat(B, 1, 1, 1)
    batch = (torch.empty(0), u_super.to(device), pos, torch.empty(B, 0, device=device), torch.empty(0), mask)
    tr = AutoregressivePushforwardTrainer(model, pde, device=device, batch_size=B, base_resolution=(T, H, W))
    return model, tr, batch, g


def check_train_step_pushforward(device, tol_fwd=1e-5, tol_grad=2e-5):
    """train_step incl. sample_windows / _labels_at / the no-grad unroll loop vs the reference's train_step (:43-163)."""
    model, tr, batch, g = _setup(device)
    seen = set()
    for seed in SEEDS:
        random.seed(seed)                                             # the reference draws from the global `random` (:82,:95)
        model.zero_grad()
        loss, pred = tr.train_step(batch, int(g["epoch"]), 0, None)
        loss.backward()
        u = int(g[f"ts{seed}_unrolled"])
        seen.add(u)
        assert abs(loss.item() - float(g[f"ts{seed}_loss"])) <= tol_fwd * abs(float(g[f"ts{seed}_loss"])), (seed, u)
        assert rel_l2(pred, g[f"ts{seed}_pred"]) < tol_fwd * (1 + u), (seed, u, rel_l2(pred, g[f"ts{seed}_pred"]))
        for k, p in model.named_parameters():
            key = f"ts{seed}_grad_{k}"
            if key in g.files:
                e = rel_l2(p.grad, g[key])
                assert e < tol_grad * (1 + u), (seed, u, k, e)
    assert seen == {0, 1, 2}


def check_test_step(device, tol=1e-5, graph=False):
    """test_step (19-window one-step losses + unrolled rollout loss) vs the reference's test_step (:165-286,:442-514)."""
    model, tr, batch, g = _setup(device)
    model.eval()
    with torch.no_grad():
        val, info = tr.test_step(batch, 0)
    assert abs(val.item() - float(g["test_loss"])) <= 10 * tol * abs(float(g["test_loss"]))
    for k, v in info.items():
        ref = float(g["test_info_" + k.replace(" ", "_").replace(",", "")])
        assert abs(float(v) - ref) <= 10 * tol * max(abs(ref), 1e-12), (k, float(v), ref)
    assert len([k for k in info if k.startswith("Step ")]) == (g["u_super"].shape[2] - 25) // 25
