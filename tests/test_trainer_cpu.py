"""CPU suite: push-forward train_step (unroll 0/1/2), test_step and the window slicer against the unmodified
reference trainer's outputs (golden) -- host-side logic, kernels replaced by the torch port."""
import pytest
import torch

from oracle.torch_port import cpu_port
from trainer_cases import check_test_step, check_train_step_pushforward


def test_train_step_pushforward_unroll_matches_reference():
    with cpu_port():
        check_train_step_pushforward("cpu")


def test_test_step_matches_reference():
    with cpu_port():
        check_test_step("cpu")


def test_create_data_distinct_steps_and_truncated_batch():
    """DataCreator.create_data (common/data_creator.py:48-78): per-sample window starts, `dp[:n]` truncation, modes."""
    from neural_pde_surrogates_b200.trainer import DataCreator
    from neural_pde_surrogates_b200 import TwoPhasePDE
    dc = DataCreator(pde=TwoPhasePDE(4, 4), time_window=5, t_resolution=40)
    u = torch.arange(3 * 1 * 40 * 2 * 2, dtype=torch.float32).reshape(3, 1, 40, 2, 2)
    steps = [5, 17, 30]
    data, labels = dc.create_data(u, steps)
    for i, s in enumerate(steps):
        assert torch.equal(data[i], u[i, :, s - 5:s]) and torch.equal(labels[i], u[i, :, s:s + 5])
    d2 = dc.create_data(u, [7, 9], mode="data")                 # fewer steps than samples: only the first two are used
    assert d2.shape[0] == 2 and torch.equal(d2[1], u[1, :, 4:9])
    l2 = dc.create_data(u, [10, 10, 10], mode="labels")         # equal steps: one strided view, no copy loop
    assert torch.equal(l2, u[:, :, 10:15])
    import pytest
    with pytest.raises(AssertionError):
        dc.create_data(u, [3])                                  # step - tw < 0
    with pytest.raises(AssertionError):
        dc.create_data(u, [38])                                 # step + tw > T


def test_epoch_loop_schedule_and_checkpoint(tmp_path):
    """train_one_epoch / train / test / save_model (trainers/base.py:219-347,472-507): batch-limit semantics (the check
    comes after the step), lr schedule every lr_step_interval epochs, checkpoints loadable with the same keys."""
    from parity_util import tiny_model
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer
    model, pde, g = tiny_model("cpu")
    T = 100
    torch.manual_seed(0)
    u = torch.rand(4, 1, T, 24, 16) * 0.5 + 0.1
    mask = (torch.rand(4, 1, 24, 16) < 0.1).float()
    pos = pde.x[None].repeat(4, 1, 1, 1)
    ds = [(torch.empty(0), u[i], pos[i], torch.empty(0), torch.empty(0), mask[i]) for i in range(4)]
    loader = torch.utils.data.DataLoader(ds, batch_size=2, shuffle=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    sched = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=[1, 2], gamma=0.4)
    tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device="cpu", batch_size=2, base_resolution=(T, 24, 16),
                                          lr_step_interval=1)
    before = {k: v.clone() for k, v in model.state_dict().items()}
    with cpu_port():
        tl, vl = tr.train(loader, num_epochs=2, valid_loader=loader, test_interval=2, lr_scheduler=sched,
                          save_path=str(tmp_path / "ckpt"), max_train_batches=0)         # still trains ONE batch per epoch
    assert len(tl) == 2 and len(vl) == 1 and all(torch.isfinite(x) for x in tl + vl)
    assert abs(opt.param_groups[0]["lr"] - 1e-3 * 0.4 * 0.4) < 1e-12                     # stepped after epochs 0 and 1
    sd = torch.load(tmp_path / "ckpt_final.pt")
    assert list(sd) == list(before) and (tmp_path / "ckpt_unrolled.pt").exists()
    assert any(not torch.equal(sd[k], before[k]) for k in sd)


def test_create_data_device_side_gather_matches_per_sample_slices():
    """DataCreator.create_data with per-sample window starts is one index op per tensor; it must equal the reference's
    Python loop of slices + cat (src/common/data_creator.py:48-78) for data, labels and both modes, equal and ragged starts."""
    from neural_pde_surrogates_b200.trainer import DataCreator
    torch.manual_seed(3)
    tw = 5
    dc = DataCreator(None, time_window=tw, t_resolution=40)
    dp = torch.randn(6, 2, 40, 4, 3)
    for steps in ([5, 9, 30, 35], [7, 7, 7], [35, 5, 20, 11, 6, 30]):
        ref_d = torch.stack([dp[i, :, s - tw:s] for i, s in enumerate(steps)])
        ref_l = torch.stack([dp[i, :, s:s + tw] for i, s in enumerate(steps)])
        d, l = dc.create_data(dp, steps)
        assert torch.equal(d, ref_d) and torch.equal(l, ref_l)
        assert torch.equal(dc.create_data(dp, steps, mode="data"), ref_d)
        assert torch.equal(dc.create_data(dp, steps, mode="labels"), ref_l)
    with pytest.raises(AssertionError):
        dc.create_data(dp, [3, 10])                                    # window would start before t = 0
