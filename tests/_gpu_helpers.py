"""Helpers shared by the GPU parity tests."""
import numpy as np
import torch

import bbo_b200 as B
from _golden import O


def impl_name(kind, eps):
    if kind == O.RBF:
        return "rbf_vmap"
    return "matern52_vmap" if eps > 0 else "matern52_cdist"


def make_modules(kind, eps, raw_ls, raw_var, const):
    raw_ls = np.asarray(raw_ls, dtype=np.float64)
    base = B.TemplateKernel(impl_name(kind, eps), ard_dim=(raw_ls.size if raw_ls.ndim == 1 else None)).double()
    with torch.no_grad():
        base.lengthscale.copy_(torch.as_tensor(raw_ls))
    cov = base
    if raw_var is not None:
        cov = B.ScaleKernel(base).double()
        with torch.no_grad():
            cov.variance.copy_(torch.tensor(float(raw_var), dtype=torch.float64))
    mean = B.ConstantMean().double()
    with torch.no_grad():
        mean.constant.copy_(torch.tensor(float(const), dtype=torch.float64))
    return mean, cov


def gp_from_params(p, X, Y, **kw):
    mean, cov = make_modules(p.kind, p.pairwise_eps, p.raw_lengthscale, p.raw_variance, p.constant)
    return B.CudaGP(torch.as_tensor(np.asarray(X)), torch.as_tensor(np.asarray(Y)).reshape(-1, 1), mean, cov,
                    noise_variance=p.noise, **kw)


def gp_from_case(c, trained=False, **kw):
    return gp_from_params(c.params(trained), c.X, c.Y, **kw)


def synth_problem(seed, n, d, kind=O.RBF, noise=1e-4, ard=True, scale=True, ls_mean=-0.2):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(2.0 * X @ w) + 0.1 * rng.standard_normal(n)
    if n > 1:
        y = (y - y.mean()) / (y.std(ddof=1) + 1e-6)
    raw_ls = rng.normal(ls_mean, 0.15, d) if ard else np.float64(ls_mean)
    p = O.GPParams(kind=kind, raw_lengthscale=raw_ls, raw_variance=(0.2 if scale else None), constant=0.05,
                   noise=noise, pairwise_eps=(1e-6 if kind == O.MATERN52 else 0.0))
    return p, X, y


def np_(t):
    return t.detach().cpu().numpy()
