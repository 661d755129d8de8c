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
        raise NotImplementedError("WarpKernel is not on the CUDA path (SURVEY section 8 row N4)")
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
        raise NotImplementedError("only ConstantMean is on the CUDA path (SURVEY section 8 row N4)")
    return mean_module.constant
