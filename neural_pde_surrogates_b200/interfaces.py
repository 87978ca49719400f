"""Data / model interface tags read by the reference trainer (src/common/interfaces.py:4-12,
src/trainers/base.py:233-241: `assert self.model.model_interface in self.model_interface`).

Enum members of two different classes never compare equal, so a processor that is dropped into the reference tree
must carry the reference's *own* `M.AR_TB` / `D.sim2d` objects.  When the reference's `common.interfaces` module is
importable (i.e. we run inside the reference tree) its enums are re-exported here unchanged;
stand-alone (this repo's trainer / bench / tests) the equivalent enums below are used."""
from enum import Enum


def _reference_enums():
    try:
        from common.interfaces import D as RD, M as RM            # the reference's src/ is on sys.path
    except Exception:                                             # noqa: BLE001  (any import problem => stand-alone)
        return None
    ok = (isinstance(RD, type) and isinstance(RM, type) and issubclass(RD, Enum) and issubclass(RM, Enum)
          and all(hasattr(RD, n) for n in ("sim1d", "sim2d", "sim1d_var_t")) and all(hasattr(RM, n) for n in ("AR_TB_GNN", "AR_TB")))
    return (RD, RM) if ok else None


_ref = _reference_enums()
if _ref is not None:
    D, M = _ref
else:
    class D(Enum):   # what one dataset element looks like
        sim1d = 0        # (c, t, x)
        sim2d = 1        # (c, t, x, y)
        sim1d_var_t = 2  # (c, t, x) with varying t

    class M(Enum):   # how the model is stepped
        AR_TB_GNN = 0    # autoregressive + temporal bundling + GNN
        AR_TB = 1        # autoregressive + temporal bundling
