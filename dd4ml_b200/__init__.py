"""dd4ml_b200 -- B200-native trust-region hot path of cruzas/DD4ML.

Drop-in optimizer classes (same names, constructor / ``step`` signatures and
hyper-parameter tables as ``dd4ml.optimizers.*``) whose vector math runs in hand-written
sm_100a CUDA kernels behind a C ABI (``include/dd4ml_b200.h``, ``csrc/``).  There is no
CPU path: constructing an optimizer on CPU parameters raises.
"""
from ._lib import DD4MLError  # noqa: F401
from .optimizers.apts_d import APTS_D  # noqa: F401
from .optimizers.apts_p import APTS_P  # noqa: F401
from .optimizers.apts_pinn import APTS_PINN  # noqa: F401
from .optimizers.lsr1 import LSR1  # noqa: F401
from .optimizers.lssr1_tr import LSSR1_TR  # noqa: F401
from .optimizers.tr import TR  # noqa: F401
from .optimizers.tradam import TRAdam  # noqa: F401
from .solvers.obs import OBS  # noqa: F401
from .install import install  # noqa: F401

__version__ = "0.1.0"
