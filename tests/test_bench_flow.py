"""CPU suite: the multi-rank control flow of bench.py (`run_legs`) on gloo with world_size 2.

Round 1's bench ran its secondary legs on rank 0 only with a data-parallel trainer and dead-locked in the gradient
all-reduce (every driver run with N > 1 crashed).  This test drives the *same* `run_legs` code with a CPU harness
(tiny model, kernels replaced by the torch port): if any rank issues a collective the other does not, the job hangs
and the test times out instead of passing."""
import json
import os
import sys
import time

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    import torch.distributed as dist
    import bench
    from neural_pde_surrogates_b200 import dp
    from oracle.torch_port import cpu_port

    class CpuHarness(bench.Harness):
        def __init__(self):
            self.dist = dist
            self.rank, self.world, self.local = dp.init_distributed("gloo")
            self.dev = torch.device("cpu")
            self.graphs = False

        def barrier(self):
            if self.world > 1:
                dist.barrier()

        def timed(self, fn, steps):
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                fn()
            ms = (time.perf_counter() - t0) * 1e3
            self.barrier()
            return self.max_over_ranks(ms)

        def pin(self, t):
            return t

        def clock_sampler(self):
            class _C:
                def __enter__(self):
                    return self

                def __exit__(self, *a):
                    return False

                def summary(self):
                    return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
            return _C()

    args = bench.make_parser().parse_args(["--gpus", str(world), "--steps", "1", "--warmup", "1", "--batch", "2",
                                           "--rollout-batch", "1", "--rollout-steps", "2", "--unroll", "2",
                                           "--no-cpu-baseline", "--no-other-configs"])
    wl = bench.Workload(H=24, W=16, width=16, modes=4, blocks=1, name="tiny")
    hx = CpuHarness()
    with cpu_port():
        line = bench.run_legs(hx, args, wl)
    if rank == 0:
        with open(os.path.join(out_dir, "line.json"), "w") as f:
            json.dump(line, f)
    else:
        assert line is None
    hx.finish()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world", [1, 2])
def test_run_legs_completes_on_every_rank(tmp_path, world):
    port = 23000 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    line = json.load(open(os.path.join(tmp_path, "line.json")))
    assert line["n_gpus"] == world and line["value"] > 0 and line["e2e"]["value"] > 0
    assert line["config"]["global_batch"] == 2 * world
    ro = line["rollout"]
    assert ro["trajectories"] == world and ro["eager"] > 0 and ro["finite"]
    assert line["train_unroll8"]["value"] > 0 and line["train_unroll8"]["unroll"] == 2
