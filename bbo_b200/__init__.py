"""bbo_b200: B200-native (sm_100a) GP-BO hot path for songlei00/bbo.

GP marginal-likelihood fit, posterior evaluation, acquisition scoring and top-k selection
as hand-written CUDA kernels behind the reference's Surrogate / AcquisitionFunction /
GradientFreeOptimizer plug-in points.  No CPU fallback.
"""
from .acquisitions import EI, PI, UCB  # noqa: F401
from .gp import CudaGP, cov_matrix  # noqa: F401
from .kernels import (ConstantMean, KumarWarp, Matern52Kernel, MLPBlock, MLPMean, MLPWarp, RBFKernel,  # noqa: F401
                      ScaleKernel, TemplateKernel, WarpKernel)
from .warped import CudaWarpedGP, surrogate  # noqa: F401

__version__ = "0.1.0"
