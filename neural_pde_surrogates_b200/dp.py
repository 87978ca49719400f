"""Data-parallel training across the GPUs of one box: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference is single-process; what sharding adds:
  1. gradients: ONE flat float32 bucket per step (complex spectral weights viewed as real pairs -- NCCL has no
     complex dtype), all-reduced with SUM;
  2. the loss `sqrt(sum_batch SE)` (autoregressivepushforwardtrainer.py:161-162) is not separable over ranks: every
     rank computes its local S_r, one scalar all-reduce gives S, and the local backward is seeded with
     1/(2 sqrt(S)); summing (not averaging) the gradients then reproduces the single-process gradient exactly;
  3. the unroll count is drawn from a generator seeded identically on all ranks (equal work), the window starts from
     a per-rank generator.
Rollout inference shards trajectories with no communication at all (`shard_trajectories`).
The same code runs on `gloo` (CPU tests, world_size 2) and `nccl`.
"""
from __future__ import annotations

import os
import random

import torch
import torch.distributed as dist


def init_distributed(backend: str | None = None):
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* set by torchrun. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


class GradBucket:
    """Flat float32 view of every parameter gradient; `.grad` tensors alias slices of one buffer, so the all-reduce
    needs no gather / scatter copies."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        sizes = [p.numel() * (2 if p.is_complex() else 1) for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views, off = [], 0
        for p, n in zip(self.params, sizes):
            seg = self.flat[off:off + n]
            v = torch.view_as_complex(seg.view(*p.shape, 2)) if p.is_complex() else seg.view(p.shape)
            self.views.append(v)
            off += n

    def attach(self):
        """Zero the bucket and point every `.grad` at its slice (call before backward)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v

    def all_reduce(self):
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def global_sqrt_loss(local_sum: torch.Tensor) -> torch.Tensor:
    """value sqrt(S) with S = sum over ranks of local_sum, gradient d local_sum / (2 sqrt(S))."""
    total = local_sum.detach().clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    root = torch.sqrt(total)
    return root + (local_sum - local_sum.detach()) / (2.0 * root)


def make_data_parallel(trainer, seed: int = 42):
    """Turn an AutoregressivePushforwardTrainer into its data-parallel version (in place)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in trainer.model.parameters():                       # identical replicas
            t = torch.view_as_real(p.data) if p.is_complex() else p.data
            dist.broadcast(t, src=0)
        from . import ops                                           # `.data` writes do not bump Parameter._version
        ops.invalidate_weight_caches()
    bucket = GradBucket(trainer.model.parameters())
    trainer.rng_unroll = random.Random(seed)                       # same on every rank
    trainer.rng_steps = random.Random(seed * 7919 + 1 + rank)      # different per rank
    trainer.loss_reduce = global_sqrt_loss
    trainer.grad_bucket = bucket

    def optimizer_step(loss):
        bucket.attach()
        loss.backward()
        bucket.all_reduce()
        trainer.optimizer.step()

    trainer.optimizer_step = optimizer_step
    return trainer


def shard_trajectories(n: int, rank: int | None = None, world: int | None = None) -> range:
    """Contiguous block of trajectory indices for this rank (rollout inference: no collective needed)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = world_size()
    per, extra = divmod(n, world)
    start = rank * per + min(rank, extra)
    return range(start, start + per + (1 if rank < extra else 0))
