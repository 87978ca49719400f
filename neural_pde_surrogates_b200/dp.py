"""Data-parallel training across the GPUs of one box: one process per GPU, NCCL over NVLink 5 / NVSwitch.

The reference is single-process; what sharding adds:
  1. gradients: one flat float32 buffer (complex spectral weights viewed as real pairs -- NCCL has no complex dtype)
     cut into ~48 MB buckets in backward order; each bucket is all-reduced (SUM) as soon as its gradients exist,
     overlapping the remaining backward pass;
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
    """Flat float32 storage of every parameter gradient, cut into buckets that are all-reduced WHILE the backward pass
    is still running.

    * `.grad` tensors alias slices of one buffer (complex spectral weights as real pairs): no gather / scatter copies;
    * the buffer is laid out in REVERSE parameter order, i.e. in the order the backward pass produces gradients, and cut
      into buckets of ~`bucket_mb`; a post-accumulate-grad hook per parameter counts a bucket down and the moment it is
      complete (and all earlier buckets have been launched: every rank issues the same collectives in the same order)
      its all-reduce is launched asynchronously on NCCL's stream, overlapping the rest of the backward pass -- the
      decoder / last U-FNO block first, the 59 MB spectral weights of each block as soon as K2's adjoint has run;
    * `finish()` launches what is left (parameters that received no gradient) and waits for all of it."""

    def __init__(self, params, bucket_mb: float = 48.0):
        self.params = [p for p in params if p.requires_grad][::-1]
        sizes = [p.numel() * (2 if p.is_complex() else 1) for p in self.params]
        dev = self.params[0].device
        self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        self.views, self.bucket_of, self.ranges = [], [], []
        off, start, limit = 0, 0, int(bucket_mb * 1e6 / 4)
        for p, n in zip(self.params, sizes):
            seg = self.flat[off:off + n]
            v = torch.view_as_complex(seg.view(*p.shape, 2)) if p.is_complex() else seg.view(p.shape)
            self.views.append(v)
            self.bucket_of.append(len(self.ranges))
            off += n
            if off - start >= limit:
                self.ranges.append((start, off))
                start = off
        if off > start:
            self.ranges.append((start, off))
        self.count = [0] * len(self.ranges)
        for b in self.bucket_of:
            self.count[b] += 1
        self.pending, self.launched, self.handles = list(self.count), 0, []
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]
        self.overlap = True

    def _make_hook(self, i):
        def hook(_param):
            if self._armed:
                self.pending[self.bucket_of[i]] -= 1
                self._launch_ready()
        return hook

    _armed = False

    def _distributed(self):
        return dist.is_initialized() and dist.get_world_size() > 1

    def _launch(self, k):
        lo, hi = self.ranges[k]
        if self._distributed():
            self.handles.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def _launch_ready(self):
        while self.launched < len(self.ranges) and self.pending[self.launched] <= 0:
            self._launch(self.launched)
            self.launched += 1

    def attach(self):
        """Zero the buffer, point every `.grad` at its slice and arm the bucket counters (call before backward)."""
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v
        self.pending, self.launched, self.handles = list(self.count), 0, []
        self._armed = self.overlap

    def finish(self):
        """Launch the buckets that did not complete during the backward pass, then wait for every all-reduce."""
        self._armed = False
        while self.launched < len(self.ranges):
            self._launch(self.launched)
            self.launched += 1
        for h in self.handles:
            h.wait()
        self.handles = []

    def all_reduce(self):                      # kept for callers of the round-1 API
        self.finish()


def global_sqrt_loss(local_sum: torch.Tensor) -> torch.Tensor:
    """value sqrt(S) with S = sum over ranks of local_sum, gradient d local_sum / (2 sqrt(S))."""
    total = local_sum.detach().clone()
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    root = torch.sqrt(total)
    return root + (local_sum - local_sum.detach()) / (2.0 * root)


def make_data_parallel(trainer, seed: int = 42, bucket_mb: float = 48.0):
    """Turn an AutoregressivePushforwardTrainer into its data-parallel version (in place)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    if dist.is_initialized() and dist.get_world_size() > 1:
        for p in trainer.model.parameters():                       # identical replicas
            t = torch.view_as_real(p.data) if p.is_complex() else p.data
            dist.broadcast(t, src=0)
        from . import ops                                           # `.data` writes do not bump Parameter._version
        ops.invalidate_weight_caches()
    bucket = GradBucket(trainer.model.parameters(), bucket_mb=bucket_mb)
    trainer.rng_unroll = random.Random(seed)                       # same on every rank
    trainer.rng_steps = random.Random(seed * 7919 + 1 + rank)      # different per rank
    trainer.loss_reduce = global_sqrt_loss
    trainer.grad_bucket = bucket

    def optimizer_step(loss):
        bucket.attach()
        loss.backward()                       # bucket all-reduces start inside (post-accumulate-grad hooks)
        bucket.finish()
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
