"""Covariance / mean modules: parameter holders with the reference's attribute names.

Mirrors bbo/algorithms/surrogates/gp/kernels.py:28-56 and means.py:19-25 so that
state dicts are interchangeable and ``CudaGP`` can also be handed the reference's own
``TemplateKernel`` / ``ScaleKernel`` / ``ConstantMean`` instances.  The arithmetic is done
by the CUDA library; ``forward`` calls it (there is no torch/CPU implementation here).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib

# impl name -> (kernel kind, PairwiseDistance eps): kernel_impl.py:40-54 (cdist) / :80-96 (vmap)
IMPLS = {
    "rbf_vmap": (_lib.KERNEL_RBF, 0.0),
    "rbf_cdist": (_lib.KERNEL_RBF, 0.0),
    "matern52_vmap": (_lib.KERNEL_MATERN52, 1e-6),
    "matern52_cdist": (_lib.KERNEL_MATERN52, 0.0),
}


class TemplateKernel(nn.Module):
    """kernels.py:28-40: raw lengthscale () or (ard_dim,), N(0,1) init, softplus transform."""

    def __init__(self, impl: str = "matern52_vmap", ard_dim: int | None = None):
        super().__init__()
        if callable(impl):
            impl = getattr(impl, "__name__", str(impl))
        if impl not in IMPLS:
            raise ValueError(f"unknown kernel impl {impl!r}; expected one of {sorted(IMPLS)}")
        self.impl = impl
        self.lengthscale = nn.Parameter(torch.randn(ard_dim) if ard_dim is not None else torch.randn(()))

    def forward(self, X1: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
        from .gp import cov_matrix
        return cov_matrix(self, X1, X2)


def RBFKernel(ard_dim: int | None = None) -> TemplateKernel:
    """kernels.py:43."""
    return TemplateKernel("rbf_vmap", ard_dim)


def Matern52Kernel(ard_dim: int | None = None) -> TemplateKernel:
    """kernels.py:44."""
    return TemplateKernel("matern52_vmap", ard_dim)


class ScaleKernel(nn.Module):
    """kernels.py:47-56: variance = softplus(raw), K = variance * base(X1, X2)."""

    def __init__(self, base_kernel: nn.Module):
        super().__init__()
        self.base_kernel = base_kernel
        self.variance = nn.Parameter(torch.randn(()))

    def forward(self, X1: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
        from .gp import cov_matrix
        return cov_matrix(self, X1, X2)


class ConstantMean(nn.Module):
    """means.py:19-25."""

    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(()))

    def forward(self, X: torch.Tensor) -> torch.Tensor:
        return self.constant.expand(*X.shape[:-1], 1)


# ---------------------------------------------------------------------------------------------
# Input warps and the learned mean.  Same constructor signatures, attribute names and
# forward arithmetic as the reference (state dicts interchange); they are tiny differentiable torch
# modules evaluated on the GPU -- the O(n^3) objective / O(n^2 q) posterior they feed is the CUDA
# library (``bbo_b200.warped.CudaWarpedGP``).
# ---------------------------------------------------------------------------------------------
def _reference_modules():
    """The reference's own warp / learned-mean / WarpKernel modules when `bbo` is importable (they are plain torch
    modules whose arithmetic this package does not replace), else None."""
    try:
        from bbo.algorithms.surrogates.gp.kernels import WarpKernel as _WK  # type: ignore
        from bbo.algorithms.surrogates.gp.means import MLPMean as _MM  # type: ignore
        from bbo.algorithms.surrogates.gp.warpers import KumarWarp as _KW, MLPBlock as _MB, MLPWarp as _MW  # type: ignore
        return _MB, _MW, _KW, _MM, _WK
    except Exception:  # noqa: BLE001 - reference absent: the structurally identical fallbacks below are used
        return None


_REF = _reference_modules()
if _REF is not None:
    MLPBlock, MLPWarp, KumarWarp, MLPMean, WarpKernel = _REF
else:
    # Fallbacks for environments without the reference: same constructor signatures, attribute names and forward
    # arithmetic as warpers.py:8-65, means.py:28-35, kernels.py:59-68, so that state dicts interchange.
    class MLPBlock(nn.Module):
        """warpers.py:8-28: Linear -> [norm] -> activation -> [dropout]."""

        def __init__(self, in_d: int, out_d: int, norm=nn.BatchNorm1d, activation=nn.ReLU, dropout: float = 0.1,
                     bias: bool = True):
            super().__init__()
            mlp = [nn.Linear(in_d, out_d, bias)]
            if norm is not None:
                mlp.append(norm(out_d))
            mlp.append(activation())
            if dropout > 0:
                mlp.append(nn.Dropout(dropout))
            self.mlp = nn.Sequential(*mlp)

        def forward(self, X: torch.Tensor) -> torch.Tensor:
            return self.mlp(X)


    class MLPWarp(nn.Module):
        """warpers.py:31-49."""

        def __init__(self, in_d: int, hidden_d, norm=None, activation=nn.Tanh, dropout: float = 0.0,
                     bias: bool = True):
            super().__init__()
            layers = []
            for h in hidden_d:
                layers.append(MLPBlock(in_d, h, norm, activation, dropout, bias))
                in_d = h
            self.layers = nn.Sequential(*layers)

        def forward(self, X: torch.Tensor) -> torch.Tensor:
            return self.layers(X)


    class KumarWarp(nn.Module):
        """warpers.py:52-65: Kumaraswamy CDF 1 - (1 - x^a)^b on inputs clipped to [1e-6, 1 - 1e-6],
        a = softplus(alpha), b = softplus(beta)."""

        def __init__(self):
            super().__init__()
            self.alpha = nn.Parameter(torch.zeros(1))
            self.beta = nn.Parameter(torch.zeros(1))
            self.eps = 1e-6

        def forward(self, X: torch.Tensor) -> torch.Tensor:
            X = X.clip(self.eps, 1 - self.eps)
            a = torch.nn.functional.softplus(self.alpha)
            b = torch.nn.functional.softplus(self.beta)
            return 1 - (1 - X.pow(a)).pow(b)


    class MLPMean(nn.Module):
        """means.py:28-35: mean(X) = Linear(in_d, 1)(warp_module(X))."""

        def __init__(self, in_d: int, warp_module: nn.Module):
            super().__init__()
            self.warp_module = warp_module
            self.mean = nn.Linear(in_d, 1)

        def forward(self, X: torch.Tensor) -> torch.Tensor:
            return self.mean(self.warp_module(X))


    class WarpKernel(nn.Module):
        """kernels.py:59-68: K(X1, X2) = base_kernel(warp(X1), warp(X2))."""

        def __init__(self, base_kernel: nn.Module, warp_module: nn.Module):
            super().__init__()
            self.warp_module = warp_module
            self.base_kernel = base_kernel

        def forward(self, X1: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
            return self.base_kernel(self.warp_module(X1), self.warp_module(X2))


def split_warps(cov_module: nn.Module):
    """Peels WarpKernel layers off a covariance module (ours or the reference's, duck-typed):
    returns (warp modules in application order, (kind, eps, raw_lengthscale, raw_variance) of the rest).
    ScaleKernel(WarpKernel(base, w)) and WarpKernel(ScaleKernel(base), w) are both accepted."""
    warps = []
    raw_var = None
    m = cov_module
    while True:
        if hasattr(m, "warp_module") and hasattr(m, "base_kernel"):
            warps.append(m.warp_module)
            m = m.base_kernel
        elif hasattr(m, "base_kernel") and hasattr(m, "variance"):
            if raw_var is not None:
                raise TypeError("nested ScaleKernels are not supported")
            raw_var = m.variance
            m = m.base_kernel
        else:
            break
    kind, eps, raw_ls, _ = describe_cov(m)
    return warps, (kind, eps, raw_ls, raw_var)


def describe_cov(cov_module: nn.Module):
    """(kind, pairwise_eps, raw_lengthscale Parameter, raw_variance Parameter | None).

    Accepts this package's modules or the reference's (duck-typed on attribute names:
    ScaleKernel has .base_kernel/.variance, TemplateKernel has .lengthscale/.impl)."""
    raw_var = None
    base = cov_module
    if hasattr(cov_module, "base_kernel") and hasattr(cov_module, "variance"):
        raw_var = cov_module.variance
        base = cov_module.base_kernel
    if hasattr(base, "warp_module"):
        raise NotImplementedError("WarpKernel needs bbo_b200.warped.CudaWarpedGP (or bbo_b200.surrogate(...))")
    if not hasattr(base, "lengthscale") or not hasattr(base, "impl"):
        raise TypeError(f"unsupported covariance module {type(cov_module).__name__}")
    impl = base.impl
    name = impl if isinstance(impl, str) else getattr(impl, "__name__", None)
    if name not in IMPLS:
        raise TypeError(f"unsupported kernel implementation {name!r}")
    kind, eps = IMPLS[name]
    return kind, eps, base.lengthscale, raw_var


def describe_mean(mean_module: nn.Module):
    if not hasattr(mean_module, "constant"):
        raise NotImplementedError("a learned mean needs bbo_b200.warped.CudaWarpedGP (or bbo_b200.surrogate(...))")
    return mean_module.constant
