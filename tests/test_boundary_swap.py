"""CPU suite: the drop-in boundary, proven inside the UNMODIFIED reference tree (skipped when no copy of the
reference is available: /root/reference or oracle/_ref).

1. `create_model` swap: with `comp.FNO/UFNO = npb.FNO/UFNO` the reference's own `models.activation_wrapper(EncProcDec)`
   builds our processors, `model.model_interface in [M.AR_TB]` holds with the REFERENCE's enum (trainers/base.py:233),
   state-dict keys / shapes / dtypes / seeded init are identical and the output matches.
2. The reference's own CLI `python -m train -C configs/train/cfg_twophase_ufno.py` (src/train.py:102-187) runs one
   epoch (sanity eval, training, validation, checkpoint, final test) on a synthetic dataset in the on-disk format of
   SURVEY.md §3.5, once with the reference's processors and once with ours: same losses.
Kernels are replaced by the torch port here (no GPU); tests/test_reference_tree_gpu.py repeats (2) on the B200."""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_loader as rl  # noqa: E402

pytestmark = pytest.mark.skipif(not rl.reference_available(), reason="no copy of the reference (run oracle/make_ref.sh)")

SWAP_CHECK = r"""
import sys, torch
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import ref_tree
ref, npb = ref_tree.prepare_reference(swap=False)
from common.interfaces import M
import models.enc_proc_dec_components as comp
from neural_pde_surrogates_b200.shell import twophase_model_kwargs
from oracle.reference_loader import twophase_pde
from oracle.torch_port import cpu_port
assert npb.M is M, "interfaces.py must re-export the reference's enums inside the reference tree"
pde = twophase_pde(ref, 24, 16)
kw = lambda: twophase_model_kwargs("UFNO", hidden_features=16, fno_modes=4, hidden_blocks=2)
torch.manual_seed(42)
m_ref = ref.models.activation_wrapper(model_class="EncProcDec", **kw(), pde=pde)
comp.FNO, comp.UFNO = npb.FNO, npb.UFNO                      # INTEGRATION.md section 2
torch.manual_seed(42)
m_new = ref.models.activation_wrapper(model_class="EncProcDec", **kw(), pde=pde)
assert type(m_new.processor[0]) is npb.UFNO
assert m_new.model_interface in [M.AR_TB], m_new.model_interface          # trainers/base.py:233
from trainers.autoregressivepushforwardtrainer import AutoregressivePushforwardTrainer as T
assert m_new.model_interface in T.model_interface and set(m_new.data_interface) & set(T.data_interface)
sa, sb = m_ref.state_dict(), m_new.state_dict()
assert list(sa) == list(sb)
for k in sa:
    assert sa[k].shape == sb[k].shape and sa[k].dtype == sb[k].dtype and torch.equal(sa[k], sb[k]), k
B = 2
u = torch.rand(B, 1, 25, 24, 16) * 0.5 + 0.1
mask = (torch.rand(B, 1, 24, 16) < 0.1).float()
args = dict(cond=torch.empty(B, 0), bc=None, pos=pde.x[None].repeat(B, 1, 1, 1), t_cond=torch.empty(B, 0), spatial_cond=mask)
y_ref = m_ref(u, **args)
with cpu_port():
    y_new = m_new(u, **args)
err = ((y_new - y_ref).norm() / y_ref.norm()).item()
assert err < 1e-5, err
print("SWAP_OK", err)
"""


def test_swap_into_reference_create_model():
    out = subprocess.run([sys.executable, "-c", SWAP_CHECK % {"root": ROOT}], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SWAP_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]


def _run_cli(workdir, swap):
    cmd = [sys.executable, os.path.join(ROOT, "tests", "ref_tree.py"), "--workdir", str(workdir), "--n", "3", "--cpu-port"]
    if not swap:
        cmd.append("--no-swap")
    cmd += ["--", "-C", "configs/train/cfg_twophase_ufno.py", "--batch_size=1", "--trainer.num_epochs=1",
            "--trainer.test_interval=1", "--model.hidden_features=16", "--model.hidden_blocks=1", "--model.fno_modes=4"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "Run Completed!" in out.stdout, out.stdout[-3000:] + out.stderr[-4000:]
    grab = lambda key: float(re.search(key + r":\s*\[?([-0-9.e+]+)", out.stdout).group(1))
    return grab("Train losses"), grab("Test loss")


@pytest.mark.timeout(1800)
def test_reference_cli_runs_one_epoch_with_swapped_processors(tmp_path):
    ref_train, ref_test = _run_cli(tmp_path, swap=False)
    new_train, new_test = _run_cli(tmp_path, swap=True)          # same dataset, same seed (configs/parse.py:318)
    assert abs(new_train - ref_train) <= 1e-4 * abs(ref_train), (new_train, ref_train)
    assert abs(new_test - ref_test) <= 1e-3 * abs(ref_test), (new_test, ref_test)
