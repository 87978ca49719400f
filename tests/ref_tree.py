"""Drive the UNMODIFIED reference tree with the B200 processors swapped in (TEST INFRASTRUCTURE).

This is the reference-side binding of INTEGRATION.md §2 applied at run time instead of by editing
`src/models/enc_proc_dec_components/__init__.py`:

    import models.enc_proc_dec_components as comp
    comp.FNO, comp.UFNO = neural_pde_surrogates_b200.FNO, neural_pde_surrogates_b200.UFNO

after which the reference's own entry point `python -m train -C configs/train/cfg_twophase_ufno.py ...`
(src/train.py:102-187) runs unchanged: its config parser, dataset, `create_model` name lookup
(models/enc_proc_dec.py:30-36), trainer interface asserts (trainers/base.py:233-241), epoch loop, evaluation
rollouts and checkpoint writer all execute the reference's code; only the processor modules are ours.

    python tests/ref_tree.py --workdir /tmp/x [--no-swap] [--cpu-port] -- -C configs/train/cfg_twophase_ufno.py --trainer.device=cuda ...

The reference sources come from /root/reference/src or the vendored copy oracle/_ref/src (oracle/make_ref.sh).
A synthetic dataset in the on-disk format of SURVEY.md §3.5 is generated under <workdir>/data/twophase/.
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_synthetic_dataset(workdir, n=4, T=501, H=96, W=64, channels=7, seed=0, n_static=0):
    """snapshots.npy [n, 7, T, H, W] (channel 6 is the one cfg_twophase_* selects, c_filter=[6]), snapshots.yaml,
    conditioning.npy [n, k], spatial_conditioning.npy [n, 1, H, W], split.yaml  (data/PDE2D.py:12-108,
    data/memmap_dataset.py:81-304)."""
    import yaml
    from numpy.lib.format import open_memmap
    d = os.path.join(workdir, "data", "twophase")
    os.makedirs(d, exist_ok=True)
    rng = np.random.default_rng(seed)
    snap = open_memmap(os.path.join(d, "snapshots.npy"), mode="w+", dtype=np.float32, shape=(n, channels, T, H, W))
    for i in range(n):                       # smooth-ish positive fields in (0.1, 0.6), like the volume fraction
        base = rng.random((1, 1, H, W), dtype=np.float32) * 0.3 + 0.15
        drift = rng.random((1, T, 1, 1), dtype=np.float32) * 0.1
        snap[i] = base + drift + rng.random((channels, T, H, W), dtype=np.float32) * 0.05
    snap.flush()
    del snap
    np.save(os.path.join(d, "conditioning.npy"), rng.random((n, max(n_static, 1))).astype(np.float32)[:, :n_static] if n_static
            else np.zeros((n, 0), dtype=np.float32))
    np.save(os.path.join(d, "spatial_conditioning.npy"), (rng.random((n, 1, H, W)) < 0.1).astype(np.float32))
    tmax, dt = 5.0, 5.0 / (T - 1)
    with open(os.path.join(d, "snapshots.yaml"), "w") as f:
        yaml.safe_dump({"x1": np.linspace(0, 1.5, H).tolist(), "x2": np.linspace(0, 1.0, W).tolist(), "tmin": 0.0,
                        "tmax": tmax + 1e-9, "dt": dt}, f)
    idx = list(range(n))
    with open(os.path.join(d, "split.yaml"), "w") as f:
        yaml.safe_dump({"train": idx[:max(n - 2, 1)], "valid": idx[-2:-1] or idx[:1], "test": idx[-1:]}, f)
    os.makedirs(os.path.join(workdir, "models"), exist_ok=True)          # utils/misc.py:44-45 wants ./models
    return d


def prepare_reference(swap=True):
    """Put the reference's src/ on sys.path (with the two import-only stubs) and swap the processors."""
    sys.path.insert(0, ROOT)
    from oracle import reference_loader as rl
    ref = rl.load_reference()                       # registers the stubs, inserts REFERENCE_SRC, imports models
    import neural_pde_surrogates_b200 as npb        # AFTER the reference: interfaces.py re-exports common.interfaces
    if swap:
        import models.enc_proc_dec_components as comp
        comp.FNO, comp.UFNO = npb.FNO, npb.UFNO     # <- the whole reference-side change
    return ref, npb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workdir", required=True)
    ap.add_argument("--no-swap", action="store_true", help="run the reference's own processors (baseline run)")
    ap.add_argument("--cpu-port", action="store_true", help="no GPU: route the swapped processors through the torch port")
    ap.add_argument("--n", type=int, default=4)
    ap.add_argument("--grid", type=int, nargs=2, default=[96, 64])
    ap.add_argument("train_argv", nargs=argparse.REMAINDER)
    a = ap.parse_args()
    argv = [x for x in a.train_argv if x != "--"]
    os.makedirs(a.workdir, exist_ok=True)
    if not os.path.exists(os.path.join(a.workdir, "data", "twophase", "snapshots.npy")):
        make_synthetic_dataset(a.workdir, n=a.n, H=a.grid[0], W=a.grid[1])
    os.chdir(a.workdir)
    ref, npb = prepare_reference(swap=not a.no_swap)
    import torch
    torch.backends.cudnn.allow_tf32 = False
    sys.argv = ["train"] + argv
    if a.cpu_port:
        from oracle.torch_port import cpu_port
        with cpu_port():
            runpy.run_module("train", run_name="__main__")
    else:
        runpy.run_module("train", run_name="__main__")


if __name__ == "__main__":
    main()
