"""-m gpu (needs >= 2 GPUs, skipped otherwise): bench.py under the driver's own multi-GPU launch line.  Round 1's
bench dead-locked for every N > 1 (rank-0-only extras issued NCCL collectives); this runs the real thing on 2 ranks."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


@pytest.mark.timeout(1500)
def test_bench_two_ranks_nccl():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "bench.py"), "--gpus", "2", "--steps", "2", "--warmup", "3",
           "--no-cpu-baseline", "--rollout-steps", "5"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1400, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["config"]["global_batch"] == 32 and line["value"] > 0
    assert line["rollout"]["n_gpus"] == 2 and line["rollout"]["cuda_graph"] > 0
    assert line["train_unroll8"]["value"] > 0
