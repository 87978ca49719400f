"""CPU suite: the data-parallel host logic on gloo with world_size 2 (the kernels are replaced by the torch port).
Checks that (sharded batch, global sqrt loss, summed flat-bucket gradients) == single-process step on the whole batch."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from neural_pde_surrogates_b200 import dp
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer
    from oracle.torch_port import cpu_port
    from parity_util import tiny_model
    dp.init_distributed("gloo")
    model, pde, g = tiny_model()
    if rank != 0:                                    # replicas start different; make_data_parallel must broadcast rank 0
        with torch.no_grad():
            for p in model.parameters():
                p.mul_(1.5)
    opt = torch.optim.SGD(model.parameters(), lr=0.0)
    tr = AutoregressivePushforwardTrainer(model, pde, optimizer=opt, device="cpu", batch_size=1, base_resolution=(501, 24, 16))
    dp.make_data_parallel(tr, seed=42, bucket_mb=0.02)           # tiny buckets: several overlapped all-reduces per step
    assert len(tr.grad_bucket.ranges) > 3
    u, labels, mask = (torch.from_numpy(g[k])[rank:rank + 1] for k in ("u", "labels", "mask"))
    pos = pde.x[None]
    with cpu_port():
        loss, _ = tr.train_step_windows(u, labels, pos, torch.empty(1, 0), mask)
        tr.optimizer_step(loss)
    unroll_draws = [tr.rng_unroll.choice(range(9)) for _ in range(4)]
    step_draws = tr.rng_steps.choices(range(100), k=4)
    torch.save({"loss": loss.detach(), "flat": tr.grad_bucket.flat.clone(), "unroll": unroll_draws, "steps": step_draws},
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_step_equals_single_process_step(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in (0, 1))
    assert torch.equal(r0["flat"], r1["flat"])                        # all-reduced gradients identical on both ranks
    assert r0["unroll"] == r1["unroll"] and r0["steps"] != r1["steps"]  # shared unroll count, per-rank window starts
    assert torch.allclose(r0["loss"], r1["loss"])

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from neural_pde_surrogates_b200 import dp
    from neural_pde_surrogates_b200.trainer import AutoregressivePushforwardTrainer
    from oracle.torch_port import cpu_port
    from parity_util import tiny_model
    model, pde, g = tiny_model()
    tr = AutoregressivePushforwardTrainer(model, pde, device="cpu", batch_size=2, base_resolution=(501, 24, 16))
    u, labels, mask = (torch.from_numpy(g[k]) for k in ("u", "labels", "mask"))
    with cpu_port():
        loss, _ = tr.train_step_windows(u, labels, pde.x[None].repeat(2, 1, 1, 1), torch.empty(2, 0), mask)
        loss.backward()
    ref = dp.GradBucket(model.parameters())
    flat = torch.cat([(torch.view_as_real(p.grad) if p.grad.is_complex() else p.grad).reshape(-1) for p in ref.params])
    assert abs(loss.item() - r0["loss"].item()) < 1e-5 * abs(loss.item())
    err = (r0["flat"] - flat).norm() / flat.norm()
    assert err < 1e-5, err


def test_shard_trajectories_partition():
    from neural_pde_surrogates_b200.dp import shard_trajectories
    for n in (0, 1, 7, 8, 64, 67):
        for world in (1, 2, 4, 8):
            parts = [list(shard_trajectories(n, r, world)) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
