"""B200-native U-FNO / FNO spectral block and rollout (drop-in for the hot path of
yoeripoels/neural-pde-surrogates).  See DESIGN.md and INTEGRATION.md."""
from .interfaces import D, M  # noqa: F401
from .proc_fno import FNO, FNO_Layer, SpectralConv2d, get_spectral_conv_with_right_spatial_dim  # noqa: F401
from .proc_ufno import UFNO  # noqa: F401
from .unet_branch import UNetModern  # noqa: F401
from .shell import (ConstrainedSurrogate, ElementWise, EncProcDec, TimeConvDense, TwoPhasePDE,  # noqa: F401
                    build_twophase_model, twophase_model_kwargs)

__version__ = "0.1.0"
