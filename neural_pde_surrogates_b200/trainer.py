"""Hot loops of the reference trainer, device-resident: push-forward training step and autoregressive rollout.

Mirror of `AutoregressivePushforwardTrainer` (reference src/trainers/autoregressivepushforwardtrainer.py):
  train_step   :43-163    u no-grad model applications, then one differentiated application, loss = sqrt(MSE_sum)
  simulate     :288-440   autoregressive rollout (same signature, same return conventions)
  test_step    :165-286 / _test_unrolled_losses :442-514
and of `DataCreator.create_data` (src/common/data_creator.py:48-78).  Differences, all on the host side:
  * windows are sliced with one view per call when all samples share the step (rollout), not a Python loop + cat;
  * `simulate(..., graph=True)` captures ONE model application in a CUDA graph and replays it per step: the state
    [B,1,tw,H,W] never leaves HBM, there is no per-step Python in the model, ~500 launches become one replay;
  * the loss reduction is pluggable so that data-parallel training reproduces the single-process gradient of the
    non-separable sqrt(sum) loss exactly (dp.py).
`process_step` (src/utils/process_output.py:8-54) is the identity for every PDE except "DIV1D" (:32,53-54), which is
outside this build, so it is a checked no-op here.
"""
from __future__ import annotations

import math
import random
from types import SimpleNamespace
from typing import List, Optional

import torch
from torch import nn

from .interfaces import M


class DataCreator:
    """Window slicer (src/common/data_creator.py:18-78, grid part only)."""

    def __init__(self, pde, time_window: int = 25, t_resolution: int = 501, **_):
        assert isinstance(time_window, int)
        self.pde, self.tw, self.t_res = pde, time_window, t_resolution

    def create_data(self, datapoints: torch.Tensor, steps: List[int], mode: str = "both"):
        assert mode in ["data", "labels", "both"]
        tw, T = self.tw, datapoints.shape[2]
        for s in steps:
            assert s - tw >= 0 and s + tw <= T, 'this step - time window combination is not valid'
        n = len(steps)
        dp = datapoints[:n]
        if all(s == steps[0] for s in steps):
            s = steps[0]
            data = dp[:, :, s - tw:s] if mode != "labels" else None
            labels = dp[:, :, s:s + tw] if mode != "data" else None
        else:
            # Device-side window gather: ONE index op per tensor instead of a Python loop of per-sample
            # slices + cat (data_creator.py:60-74).  rows[i, j] = steps[i] - tw + j for the data, steps[i] + j for the labels.
            st = torch.as_tensor(steps, device=dp.device, dtype=torch.long)
            off = torch.arange(tw, device=dp.device, dtype=torch.long)
            bi = torch.arange(n, device=dp.device, dtype=torch.long)[:, None]

            def window(start):                                   # dp [n, C, T, H, W] -> [n, C, tw, H, W]
                return dp[bi, :, (start[:, None] + off)[:, :]].transpose(1, 2).contiguous()
            data = window(st - tw) if mode != "labels" else None
            labels = window(st) if mode != "data" else None
        if mode == "data":
            return data
        if mode == "labels":
            return labels
        return data, labels


def _process_step_identity(pde):
    if f"{pde}" == "DIV1D":
        raise NotImplementedError("boundary / minimum clamps of the DIV1D dataset are outside the B200 build")


class GraphedModelStep:
    """One model application captured in a CUDA graph: out = model(u, cond, pos, spatial_cond).  Every input lives in a
    static buffer owned by this object and is refreshed on each call, so a replay can never see stale conditioning."""

    def __init__(self, model, u, cond, pos, spatial_cond, warmup: int = 3):
        own = lambda t: None if t is None else t.detach().clone()
        self.u = own(u)
        self.static = dict(cond=own(cond), pos=own(pos), spatial_cond=own(spatial_cond))
        self.kw = dict(cond=self.static["cond"], bc=None, pos=self.static["pos"], t_cond=None,
                       spatial_cond=self.static["spatial_cond"])
        side = torch.cuda.Stream(device=u.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                 # fills twiddle-table / cuDNN plan caches outside the capture
                model(self.u, **self.kw)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = model(self.u, **self.kw)

    def signature(self, u, cond, pos, spatial_cond):
        sig = lambda t: None if t is None else (tuple(t.shape), t.dtype, t.device)
        return (sig(u), sig(cond), sig(pos), sig(spatial_cond))

    def matches(self, u, cond, pos, spatial_cond) -> bool:
        return self.signature(u, cond, pos, spatial_cond) == self.signature(self.u, *(self.static[k] for k in ("cond", "pos", "spatial_cond")))

    def __call__(self, u, cond=None, pos=None, spatial_cond=None, refresh: bool = True):
        if refresh:
            for k, t in (("cond", cond), ("pos", pos), ("spatial_cond", spatial_cond)):
                buf = self.static[k]
                if buf is not None and t is not None and buf.numel() and t.data_ptr() != buf.data_ptr():
                    buf.copy_(t)
        if u.data_ptr() != self.u.data_ptr():
            self.u.copy_(u)
        self.graph.replay()
        return self.out


class GraphedRollout:
    """K consecutive model applications captured in ONE CUDA graph: the state [B,1,tw,H,W] is handed from one
    application to the next inside the graph (no host round trip, no per-step copy); the K predictions stay in static
    buffers and are copied out once per replay.  Replaying ceil(n / K) times gives an n-step rollout."""

    def __init__(self, model, u, cond, pos, spatial_cond, steps_per_graph: int, warmup: int = 2):
        own = lambda t: None if t is None else t.detach().clone()
        self.K = int(steps_per_graph)
        self.u = own(u)
        self.static = dict(cond=own(cond), pos=own(pos), spatial_cond=own(spatial_cond))
        self.kw = dict(cond=self.static["cond"], bc=None, pos=self.static["pos"], t_cond=None,
                       spatial_cond=self.static["spatial_cond"])
        side = torch.cuda.Stream(device=u.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model(self.u, **self.kw)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            outs, state = [], self.u
            for _ in range(self.K):
                state = model(state, **self.kw)
                outs.append(state)
            self.outs = outs

    def matches(self, u, cond, pos, spatial_cond) -> bool:
        sig = lambda t: None if t is None else (tuple(t.shape), t.dtype, t.device)
        return (sig(u), sig(cond), sig(pos), sig(spatial_cond)) == \
            (sig(self.u), *(sig(self.static[k]) for k in ("cond", "pos", "spatial_cond")))

    def __call__(self, u, cond=None, pos=None, spatial_cond=None):
        """Returns the K predictions that follow `u` (views of static buffers: clone before the next replay)."""
        for k, t in (("cond", cond), ("pos", pos), ("spatial_cond", spatial_cond)):
            buf = self.static[k]
            if buf is not None and t is not None and buf.numel() and t.data_ptr() != buf.data_ptr():
                buf.copy_(t)
        if u.data_ptr() != self.u.data_ptr():
            self.u.copy_(u)
        self.graph.replay()
        return self.outs


class GraphedTrainStep:
    """One whole optimizer step -- push-forward applications, differentiated application, loss, backward, gradient
    all-reduce, Adam -- captured in ONE CUDA graph and replayed with new windows copied into static buffers.
    At the reference's CPU-runnable batch (4) an eager step is ~1800 kernel launches and host-launch-bound; a replay is
    one launch.  Needs a capturable optimizer (`torch.optim.Adam(..., capturable=True)`) and fixed shapes / unroll count
    (one graph per unroll count).  The packed tensor-core operands of the weights are rebuilt inside the graph on every
    replay, so the weights the replay uses are always the current ones."""

    def __init__(self, trainer, data, labels, x, conditioning, spatial_conditioning, unrolled: int = 0, warmup: int = 3):
        own = lambda t: None if t is None else t.detach().clone()
        self.tr, self.unrolled = trainer, int(unrolled)
        self.data, self.labels = own(data), own(labels)
        self.next_labels = [own(labels) for _ in range(self.unrolled)]        # labels of the windows after each unroll
        self.x, self.cond, self.sc = own(x), own(conditioning), own(spatial_conditioning)
        for group in trainer.optimizer.param_groups:
            if not group.get("capturable", False):
                raise ValueError("GraphedTrainStep needs a capturable optimizer, e.g. torch.optim.Adam(params, capturable=True)")
        side = torch.cuda.Stream(device=data.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):                                           # cuDNN autotuning, lazy state, table caches
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        from . import ops
        before = ops.counters()["launches"]
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        self.launches_per_replay = ops.counters()["launches"] - before       # our own kernels inside one replay

    def _step(self):
        nl = (lambda k: self.next_labels[k - 1]) if self.unrolled else None
        loss, _ = self.tr.train_step_windows(self.data, self.labels if not self.unrolled else self.next_labels[-1], self.x,
                                             self.cond, self.sc, unrolled=self.unrolled, next_labels=nl)
        self.tr.optimizer_step(loss)
        return loss.detach()

    def __call__(self, data, labels, next_labels=None):
        """Copy the new windows in, replay, return the (static) loss tensor."""
        self.data.copy_(data, non_blocking=True)
        if self.unrolled:
            for k in range(self.unrolled):
                self.next_labels[k].copy_(next_labels(k + 1), non_blocking=True)
        else:
            self.labels.copy_(labels, non_blocking=True)
        self.graph.replay()
        from . import ops
        ops.add_launches(self.launches_per_replay)
        return self.loss


class AutoregressivePushforwardTrainer:
    model_interface = [M.AR_TB]

    def __init__(self, model: nn.Module, pde, criterion=None, optimizer=None, device="cuda", batch_size: int = 16,
                 time_window: int = 25, base_resolution=(501, 96, 64), unrolling: int = 8, lr_step_interval: int = 25,
                 nr_gt_steps: int = 1, process_settings: Optional[dict] = None, rng_unroll=None, rng_steps=None,
                 loss_reduce=None, grad_sync=None):
        self.model, self.optimizer = model, optimizer
        self.criterion = criterion if criterion is not None else nn.MSELoss(reduction="sum")   # defaults/criterion.py:5-8
        self.data = SimpleNamespace(pde=pde)
        self.config = SimpleNamespace(device=device, batch_size=batch_size, time_window=time_window,
                                      base_resolution=tuple(base_resolution), unrolling=unrolling,
                                      lr_step_interval=lr_step_interval, nr_gt_steps=nr_gt_steps,
                                      process_settings=process_settings or {})
        self.data_creator = DataCreator(pde=pde, time_window=time_window, t_resolution=base_resolution[0])
        # the reference draws both from the global `random` module (:82,:95); dp.py passes a shared generator for the
        # unroll count (equal work on all ranks) and a per-rank generator for the window starts
        self.rng_unroll = rng_unroll if rng_unroll is not None else random
        self.rng_steps = rng_steps if rng_steps is not None else random
        self.loss_reduce = loss_reduce if loss_reduce is not None else torch.sqrt              # :161-162
        self.grad_sync = grad_sync
        self._graphs = {}

    # ------------------------------------------------------------------------------------------ training
    def sample_windows(self, epoch: int, t_res: Optional[int] = None):
        """(unrolled_graphs, random_steps) exactly as train_step draws them (:78-95)."""
        t_res = self.data_creator.t_res if t_res is None else t_res
        tw = self.data_creator.tw
        max_unrolling = min(epoch // self.config.lr_step_interval, self.config.unrolling)
        unrolled = self.rng_unroll.choice(list(range(max_unrolling + 1)))
        steps = list(range(tw, t_res - tw - tw * unrolled + 1))
        return unrolled, self.rng_steps.choices(steps, k=self.config.batch_size)

    def train_step(self, batch, epoch: int, batch_idx: int = 0, loader=None):
        """Push-forward step on a batch (u_base, u_super, x, conditioning, t_conditioning, spatial_conditioning) of
        whole trajectories.  Returns (loss, pred) like the reference (:43-163)."""
        _, u_super, x, conditioning, t_conditioning, spatial_conditioning = batch
        if torch.numel(t_conditioning) != 0:
            raise NotImplementedError("time-varying conditioning is not used by the twophase configs")
        if torch.numel(spatial_conditioning) == 0:
            spatial_conditioning = None
        unrolled, random_steps = self.sample_windows(epoch)
        data, labels = self.data_creator.create_data(u_super, random_steps)
        dev = self.config.device
        data, labels = data.to(dev, non_blocking=True), labels.to(dev, non_blocking=True)
        return self.train_step_windows(data, labels, x, conditioning, spatial_conditioning, unrolled=unrolled,
                                       next_labels=lambda k: self._labels_at(u_super, random_steps, k).to(dev, non_blocking=True))

    def _labels_at(self, u_super, random_steps, k):
        tw = self.data_creator.tw
        return self.data_creator.create_data(u_super, [s + k * tw for s in random_steps], mode="labels")

    def train_step_windows(self, data, labels, x, conditioning, spatial_conditioning, unrolled: int = 0, next_labels=None):
        """The differentiable part of train_step on already-sliced windows (:115-163)."""
        _process_step_identity(self.data.pde)
        kw = dict(cond=conditioning, bc=None, pos=x, t_cond=None, spatial_cond=spatial_conditioning)
        with torch.no_grad():
            for k in range(unrolled):                              # push-forward: predictions fed back in, no grad
                data = self.model(data, **kw)
                labels = next_labels(k + 1)
        pred = self.model(data, **kw)
        loss = self.loss_reduce(self.criterion(pred, labels))
        return loss, pred

    def optimizer_step(self, loss):
        """zero_grad -> backward -> [gradient all-reduce] -> step  (src/trainers/base.py:490-493)."""
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.grad_sync is not None:
            self.grad_sync()
        self.optimizer.step()

    # ------------------------------------------------------------------------------------------ epoch loop
    def train_one_epoch(self, loader, epoch: int, lr_scheduler=None, max_train_batches=float("inf")):
        """One pass over `loader` (batches of whole trajectories, as PDE2DDataset yields them): H2D, zero_grad ->
        train_step -> backward -> [gradient all-reduce] -> step, loss accounting and the lr schedule exactly as
        TrainInterface.train_one_epoch (trainers/base.py:472-507; the batch-limit check comes AFTER the step, :498)."""
        self.model.train()
        dev = self.config.device
        total = 0
        for batch_idx, batch in enumerate(loader):
            batch = tuple(t.to(dev, non_blocking=True) if isinstance(t, torch.Tensor) else t for t in batch)
            loss, _ = self.train_step(batch, epoch, batch_idx, loader=loader)
            self.optimizer_step(loss)
            total = total + loss.detach() / batch[1].shape[0]                  # utils.get_batch_size
            if batch_idx >= max_train_batches:
                break
        total = total / len(loader)
        if lr_scheduler is not None and (epoch + 1) % self.config.lr_step_interval == 0:   # :504-506
            lr_scheduler.step()
        return total

    def train(self, train_loader, num_epochs: int, valid_loader=None, test_interval: int = 25, lr_scheduler=None,
              save_path: Optional[str] = None, max_train_batches=float("inf")):
        """Epoch loop of TrainInterface.train (trainers/base.py:219-347) without the logging / W&B plumbing: training
        epochs, validation with test_step every `test_interval` epochs, best / final checkpoints with the reference's
        state_dict keys (`torch.save(model.state_dict())`, :349-355).  Returns (train_losses, val_losses)."""
        train_losses, val_losses, best = [], [], float("inf")
        for epoch in range(num_epochs):
            train_losses.append(self.train_one_epoch(train_loader, epoch, lr_scheduler, max_train_batches))
            if valid_loader is not None and (epoch + 1) % test_interval == 0:
                val = self.test(valid_loader)
                val_losses.append(val)
                if save_path is not None and float(val) < best:
                    best = float(val)
                    self.save_model(save_path + "_unrolled.pt")
        if save_path is not None:
            self.save_model(save_path + "_final.pt")
        return train_losses, val_losses

    def test(self, loader, max_test_batches=float("inf")):
        """Mean of test_step's primary loss over the loader (trainers/base.py:378-470, grid path)."""
        self.model.eval()
        dev = self.config.device
        losses = []
        with torch.no_grad():
            for batch_idx, batch in enumerate(loader):
                batch = tuple(t.to(dev) if isinstance(t, torch.Tensor) else t for t in batch)
                losses.append(self.test_step(batch, batch_idx)[0])
                if batch_idx >= max_test_batches:
                    break
        return torch.mean(torch.stack(losses))

    def save_model(self, path: str):
        torch.save(self.model.state_dict(), path)                              # trainers/base.py:349-355

    # ------------------------------------------------------------------------------------------ rollout
    def graphed_train_step(self, data, labels, x, conditioning, spatial_conditioning, unrolled: int = 0):
        """The CUDA-graph version of train_step_windows + optimizer_step for these shapes (built on first use)."""
        key = ("train", tuple(data.shape), int(unrolled), data.device.index)
        g = self._graphs.get(key)
        if g is None:
            g = GraphedTrainStep(self, data, labels, x, conditioning, spatial_conditioning, unrolled)
            self._graphs[key] = g
        return g

    def _model_step(self, pred, conditioning, x, spatial_cond, graph: bool):
        if not graph:
            return self.model(pred, cond=conditioning, bc=None, pos=x, t_cond=None, spatial_cond=spatial_cond)
        key = (tuple(pred.shape), pred.device.index)
        g = self._graphs.get(key)
        if g is None or not g.matches(pred, conditioning, x, spatial_cond):
            g = GraphedModelStep(self.model, pred, conditioning, x, spatial_cond)
            self._graphs[key] = g
        return g(pred, conditioning, x, spatial_cond).clone()

    def simulate(self, u, conditioning, x, compute_loss, include_data, nr_gt_steps, t_res,
                 t_conditioning=torch.empty(0), spatial_conditioning=torch.empty(0), clip_min=True, use_bc=True,
                 u_bc=None, u_mask=None, divide_by_t=True, graph: bool = False, steps_per_graph: int = 1):
        """Autoregressive rollout with the reference's signature and return conventions (:288-440).
        Returns: losses | data_pred | (losses, (data_gt, data_pred)) depending on compute_loss / include_data."""
        tw = self.data_creator.tw
        use_mask = u_mask is not None
        if compute_loss is False and use_mask:
            raise ValueError("Mask supplied for computing the loss, but 'compute_loss'=False!")
        if compute_loss is True and u.shape[2] < t_res:
            raise ValueError("Cannot compute loss if no ground-truth simulation is provided for the full rollout")
        if u_bc is None:
            u_bc = u
        if use_bc and u_bc.shape[2] < t_res and f"{self.data.pde}" == "DIV1D":
            raise ValueError("Cannot set BCs if the provided BC information is <= the unrolling time")
        if u.shape[2] < nr_gt_steps * tw:
            raise ValueError(f"The training data is shorter than the specified number of unrolling steps: "
                             f"{nr_gt_steps} * {tw} = {nr_gt_steps * tw}, but u.shape[2] = {u.shape[2]}")
        if torch.numel(t_conditioning) != 0:
            raise NotImplementedError("time-varying conditioning is not used by the twophase configs")
        _process_step_identity(self.data.pde)
        spatial_cond = spatial_conditioning if torch.numel(spatial_conditioning) != 0 else None
        batch_size, dev = u.shape[0], self.config.device
        pred = u[:, :, tw * nr_gt_steps - tw: tw * nr_gt_steps].to(dev)        # first input = ground truth (:332-334)
        data_gt, data_pred, losses = [pred], [pred], []
        n_t = 0
        npix = math.prod(self.config.base_resolution[1:])
        steps = list(range(tw * nr_gt_steps, t_res - tw + 1, tw))
        chain, queue = None, []
        if graph and steps_per_graph > 1 and len(steps) >= steps_per_graph:
            # K model applications per graph replay: the state is handed on inside the graph (GraphedRollout)
            key = ("rollout", tuple(pred.shape), int(steps_per_graph), pred.device.index)
            chain = self._graphs.get(key)
            if chain is None or not chain.matches(pred, conditioning, x, spatial_cond):
                chain = GraphedRollout(self.model, pred, conditioning, x, spatial_cond, steps_per_graph)
                self._graphs[key] = chain
        for i, step in enumerate(steps):
            labels = u[:, :, step:step + tw].to(dev) if compute_loss else None
            if chain is not None and (queue or len(steps) - i >= chain.K):
                if not queue:
                    queue = [o.clone() for o in chain(pred, conditioning, x, spatial_cond)]
                pred = queue.pop(0)
            else:
                pred = self._model_step(pred, conditioning, x, spatial_cond, graph)
            if compute_loss and use_mask:
                lm = u_mask[:, :, step:step + tw].to(dev)
                pred, labels = pred * lm, labels * lm
            if compute_loss:
                losses.append(self.criterion(pred, labels) / npix / batch_size)
                data_gt.append(labels)
            data_pred.append(pred)
            n_t += tw
        if divide_by_t and compute_loss:
            losses = [v / n_t for v in losses]
        if compute_loss and not include_data:
            return losses
        if not compute_loss and include_data:
            return data_pred
        return losses, (data_gt, data_pred)

    def _test_unrolled_losses(self, batch, include_data=False, max_test_len=None, divide_by_t=True):
        """Full-trajectory unrolled loss (+ the baseline-solver loss when u_base is given) (:442-514)."""
        u_base, u_super, x, conditioning, t_conditioning, spatial_conditioning = batch
        t_res = self.data_creator.t_res
        out = self.simulate(u_super, conditioning, x, t_conditioning=t_conditioning,
                            spatial_conditioning=spatial_conditioning, compute_loss=True, include_data=include_data,
                            nr_gt_steps=self.config.nr_gt_steps, t_res=t_res, divide_by_t=divide_by_t)
        losses_tmp, sims = (out if include_data else (out, None))
        tw, bs = self.data_creator.tw, u_super.shape[0]
        base, n_t = [], 0
        for step in range(tw * self.config.nr_gt_steps, t_res - tw + 1, tw):
            if torch.numel(u_base) == 0:
                base.append(torch.Tensor(0))
                continue
            ls = u_super[:, :, step:step + tw]
            lb = u_base[:, :, step:step + tw]
            base.append(self.criterion(ls, lb) / math.prod(self.config.base_resolution[1:]) / bs)
            n_t += tw
        base_loss = torch.sum(torch.stack(base))
        if divide_by_t:
            base_loss = base_loss / (n_t if n_t > 0 else 1)
        total = torch.sum(torch.stack(losses_tmp))
        if include_data:
            gt, pr = sims
            return total, base_loss, [torch.cat(gt, dim=2), torch.cat(pr, dim=2), [{} for _ in range(bs)]]
        return total, base_loss

    def test_step(self, batch, batch_idx: int = 0, include_data=False):
        """One-step losses at every window + the unrolled loss (:165-286)."""
        _, u_super, x, conditioning, t_conditioning, spatial_conditioning = batch
        t_res, tw, dev = self.data_creator.t_res, self.data_creator.tw, self.config.device
        bs = u_super.shape[0]
        sc = spatial_conditioning if torch.numel(spatial_conditioning) != 0 else None
        losses, per_step = [], {}
        for step in range(tw, t_res - tw + 1, tw):
            data, labels = self.data_creator.create_data(u_super, [step] * bs)
            data, labels = data.to(dev), labels.to(dev)
            pred = self.model(data, cond=conditioning, bc=None, pos=x, t_cond=None, spatial_cond=sc)
            l = self.criterion(pred, labels) / bs
            losses.append(l)
            per_step[f"Step {step}, mean loss"] = l
        losses = torch.stack(losses)
        un = self._test_unrolled_losses(batch, include_data)
        info = {"Unrolled base losses": un[1], "Unrolled forward losses": un[0], "Mean per-step loss": torch.mean(losses),
                **per_step}
        if include_data:
            return torch.mean(un[0]), info, un[2]
        return torch.mean(un[0]), info
