"""Data / model interface tags read by the reference trainer (src/common/interfaces.py:4-12,
src/trainers/base.py:233-241).  Members compare by name and value so that a model built from these classes passes
the reference trainer's interface checks when dropped into its tree."""
from enum import Enum


class D(Enum):   # what one dataset element looks like
    sim1d = 0        # (c, t, x)
    sim2d = 1        # (c, t, x, y)
    sim1d_var_t = 2  # (c, t, x) with varying t


class M(Enum):   # how the model is stepped
    AR_TB_GNN = 0    # autoregressive + temporal bundling + GNN
    AR_TB = 1        # autoregressive + temporal bundling
