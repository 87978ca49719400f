"""-m gpu: the reference's OWN entry point on the B200 with our processors swapped in (the drop-in boundary).

`python -m train -C configs/train/cfg_twophase_ufno.py --trainer.device=cuda ...` (src/train.py:102-187) is run by
tests/ref_tree.py from the vendored copy oracle/_ref/src (made by oracle/make_ref.sh; /root/reference does not exist
on the GPU box) on a synthetic dataset in the on-disk format of SURVEY.md §3.5 -- sanity evaluation, one training
epoch, validation, checkpoint and the final test all execute the reference's code at the FULL config (width 192,
3 blocks, modes 10).  Run twice: with the reference's processors (cuFFT + cuBLAS + cuDNN, TF32 off) and with ours;
the reported losses must agree."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rl.reference_available(), reason="oracle/_ref missing (oracle/make_ref.sh)")]


def _run(workdir, swap, cfg):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_tree.py"), "--workdir", str(workdir), "--n", "4"]
    if not swap:
        cmd.append("--no-swap")
    cmd += ["--", "-C", cfg, "--trainer.device=cuda", "--batch_size=2", "--trainer.num_epochs=1", "--trainer.test_interval=1"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert out.returncode == 0 and "Run Completed!" in out.stdout, out.stdout[-3000:] + out.stderr[-4000:]
    grab = lambda key: float(re.search(key + r":\s*\[?([-0-9.e+]+)", out.stdout).group(1))
    return grab("Train losses"), grab("Test loss"), out.stdout


@pytest.mark.timeout(3600)
@pytest.mark.parametrize("cfg", ["configs/train/cfg_twophase_ufno.py", "configs/train/cfg_twophase_ufno_fno.py"])
def test_reference_train_cli_on_gpu_with_b200_processors(tmp_path, cfg):
    ref_train, ref_test, _ = _run(tmp_path, False, cfg)
    new_train, new_test, log = _run(tmp_path, True, cfg)
    assert "Loaded device: cuda" in log
    assert abs(new_train - ref_train) <= 1e-4 * abs(ref_train), (new_train, ref_train)
    assert abs(new_test - ref_test) <= 1e-3 * abs(ref_test), (new_test, ref_test)
