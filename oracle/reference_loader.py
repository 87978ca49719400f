"""Import the *unmodified* reference modules (TEST INFRASTRUCTURE ONLY).

Two locations, tried in this order:
  1. /root/reference/src  -- the read-only mount of the build container (absent on the GPU box);
  2. oracle/_ref/src      -- the verbatim copy made by `oracle/make_ref.sh` (git-ignored, travels to the GPU box).
Used by `tests/golden/make_golden.py` (fixture generator), by the tests that pin the oracle / torch port / product
modules against the real reference, and by the `cpu_baseline` / `--impl reference` legs of bench.py (which time the
reference's own CPU implementation).  The product package never imports this file.

Two import-only stubs are needed: `torch_geometric.data.Data` and
`mmap_ninja.ragged.RaggedMmap`; neither symbol is touched on the grid path.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("PDES_REFERENCE_SRC"), "/root/reference/src", os.path.join(_HERE, "_ref", "src")]


def _find_src():
    for c in _CANDIDATES:
        if c and os.path.isdir(os.path.join(c, "models")):
            return c
    return None


REFERENCE_SRC = _find_src() or "/root/reference/src"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_SRC, "models"))


def reference_kind() -> str:
    """'mount' (read-only /root/reference), 'vendored' (oracle/_ref copy) or 'absent'."""
    if not reference_available():
        return "absent"
    return "vendored" if os.path.realpath(REFERENCE_SRC).startswith(os.path.realpath(_HERE)) else "mount"


def load_reference():
    """Returns a namespace with the reference's modules (models, pdes, proc_fno, proc_ufno, ...)."""
    if not reference_available():
        raise RuntimeError(f"reference sources not found under {REFERENCE_SRC}")
    for pkg, sub, sym in (("torch_geometric", "data", "Data"), ("mmap_ninja", "ragged", "RaggedMmap")):
        if pkg not in sys.modules:
            p, s = types.ModuleType(pkg), types.ModuleType(f"{pkg}.{sub}")
            setattr(s, sym, type(sym, (), {}))
            setattr(p, sub, s)
            sys.modules[pkg], sys.modules[f"{pkg}.{sub}"] = p, s
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):        # silences the torch_cluster notice
        import models  # noqa: F401
        import pdes
        from models.enc_proc_dec_components import proc_fno, proc_ufno, proc_unet_modern
        from trainers import autoregressivepushforwardtrainer as ref_trainer
    ns = types.SimpleNamespace(models=models, pdes=pdes, proc_fno=proc_fno, proc_ufno=proc_ufno,
                               proc_unet_modern=proc_unet_modern, trainer=ref_trainer)
    return ns


def twophase_pde(ref, H=96, W=64, n_cond_static=0):
    """PDE2D metadata as PDE2DDataset builds it for the twophase experiment (src/data/PDE2D.py:73-90,
    src/pdes/base.py:34-52)."""
    return ref.pdes.PDE2D(tmin=0.0, tmax=5.0, nt=501, L1=1.5, L2=1.0, nx1=H, nx2=W, x=None, name="twophase",
                          n_cond_static=n_cond_static, n_cond_dynamic=0, n_cond_spatial=1)
