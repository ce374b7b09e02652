"""gpflowopt_b200 -- B200-native batched acquisition evaluation behind GPflowOpt's acquisition/optimizer API.

Only the data-parallel hot path of GPflowOpt lives here: scoring candidate sets with
EI / PoI / PoF / LCB, their sums and products and HVPoI, on hand-written FP64 CUDA kernels for sm_100a
(libgpfo.so, C-ABI in include/gpfo.h).  There is no CPU fallback.
"""
from . import acquisition, bo, design, domain, optim, pareto, scaling, transforms
from .bo import BayesianOptimizer
from ._lib import GpfoError, NotPositiveDefiniteError
from .engine import get_device, set_device
from .gpr import GPR, RBF, Matern32, Matern52

__version__ = "0.1.0"
