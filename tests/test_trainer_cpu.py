"""CPU suite: push-forward train_step (unroll 0/1/2), test_step and the window slicer against the unmodified
reference trainer's outputs (golden) -- host-side logic, kernels replaced by the torch port."""
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
