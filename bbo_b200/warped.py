"""CudaWarpedGP: the reference GP (gp/gp.py:30-83) with an input warp and/or a learned mean
 -- ``WarpKernel`` (gp/kernels.py:59-68) around ``KumarWarp`` / ``MLPWarp``
(gp/warpers.py:31-65), and ``MLPMean`` (gp/means.py:28-35).

Split of the work:
  * the warp / mean modules are tiny differentiable torch modules on the GPU: O(n d h) flops;
  * everything cubic or quadratic in n -- covariance, Cholesky, objective value, its gradient w.r.t. the
    kernel hyper-parameters AND w.r.t. the warped inputs and the residual targets
    (``bbo_gp_fit_value_grad_inputs``), the posterior and the pool scoring -- is the CUDA library.
``objective()`` returns the reference's objective (objectives.py:34-45, = +log p(y|X)) as a
differentiable tensor, so ``train()`` is literally gp.py:54-63: Adam over ``mean.parameters() +
cov.parameters()``.  After training, an inner ``CudaGP`` on (warp(X), Y - mean(X)) serves predict /
score_pool; the pool prologue is ``bbo_warp_kumaraswamy`` for KumarWarp and the module itself otherwise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib
from .abstractions import Surrogate
from .gp import CudaGP, _as_2d_f64, _check_2d, _hyper, _raise_not_pd
from .kernels import ConstantMean, Matern52Kernel, ScaleKernel, TemplateKernel, split_warps

_F64 = torch.float64


class _Objective(torch.autograd.Function):
    """objectives.py:34-45 on the device with a hand-written backward: value (1,1) from
    (Xw (n,dw), resid (n,), raw_lengthscale, raw_variance | None)."""

    @staticmethod
    def forward(ctx, owner, Xw, resid, raw_ls, raw_var):
        lib, dev = owner._lib, owner._device
        n, dw = Xw.shape
        ard = raw_ls.dim() == 1
        if ard and raw_ls.numel() != dw:
            raise ValueError(f"ARD lengthscale has {raw_ls.numel()} dims, the warped inputs have {dw}")
        with torch.cuda.device(dev):
            xw = Xw.detach().to(_F64).contiguous()
            r = resid.detach().to(_F64).reshape(-1).contiguous()
            parts = [torch.zeros(1, dtype=_F64, device=dev), raw_ls.detach().reshape(-1).to(_F64)]
            if raw_var is not None:
                parts.append(raw_var.detach().reshape(1).to(_F64))
            theta = torch.cat(parts).contiguous()
            hyp = torch.empty(2 * dw + 2, dtype=_F64, device=dev)
            st = _lib.stream_ptr(dev)
            _lib.check(lib.bbo_gp_theta_to_hyp(dw, 1, int(ard), int(raw_var is not None), _lib.ptr(theta),
                                               _lib.ptr(hyp), st))
            h = _hyper(owner._kind, dw, owner._eps, owner._noise_variance)
            ws = owner._workspace("fit", int(lib.bbo_gp_fit_workspace_bytes(n, dw, 1)))
            value = torch.empty(1, dtype=_F64, device=dev)
            grad = torch.empty(dw + 2, dtype=_F64, device=dev)
            gX = torch.empty(n, dw, dtype=_F64, device=dev)
            gy = torch.empty(n, dtype=_F64, device=dev)
            info = torch.zeros(1, dtype=torch.int32, device=dev)
            _lib.check(lib.bbo_gp_fit_value_grad_inputs(
                C.byref(h), n, _lib.ptr(xw), _lib.ptr(r), _lib.ptr(hyp), _lib.ptr(value), _lib.ptr(grad),
                _lib.ptr(gX), _lib.ptr(gy), _lib.ptr(info), _lib.ptr(ws), ws.numel(), st),
                "bbo_gp_fit_value_grad_inputs")
            i = int(info.item())
            if i != 0:
                _raise_not_pd(i)
            rl = raw_ls.detach().to(_F64)
            sg = torch.where(rl > 20.0, torch.ones_like(rl), torch.sigmoid(rl))      # softplus'
            g_ls = grad[:dw] * sg if ard else grad[:dw].sum() * sg
            g_var = None
            if raw_var is not None:
                rv = raw_var.detach().to(_F64)
                g_var = grad[dw] * torch.where(rv > 20.0, torch.ones_like(rv), torch.sigmoid(rv))
        ctx.save_for_backward(gX, gy, g_ls, g_var if g_var is not None else torch.zeros((), device=dev))
        ctx.has_var = raw_var is not None
        ctx.meta = (Xw.dtype, resid.dtype, resid.shape, raw_ls.dtype, raw_ls.shape,
                    None if raw_var is None else (raw_var.dtype, raw_var.shape))
        return value.reshape(1, 1).to(Xw.dtype)

    @staticmethod
    def backward(ctx, go):
        gX, gy, g_ls, g_var = ctx.saved_tensors
        g = go.reshape(()).to(_F64)
        xdt, rdt, rshape, ldt, lshape, vmeta = ctx.meta
        out_var = None
        if ctx.has_var:
            out_var = (g * g_var).to(vmeta[0]).reshape(vmeta[1])
        return (None, (g * gX).to(xdt), (g * gy).to(rdt).reshape(rshape), (g * g_ls).to(ldt).reshape(lshape),
                out_var)


class CudaWarpedGP(Surrogate):
    def __init__(self, X: torch.Tensor, Y: torch.Tensor, mean_module: Optional[nn.Module] = None,
                 cov_module: Optional[nn.Module] = None, *, device="cuda", lr: float = 0.01, epochs: int = 100,
                 noise_variance: float = 1e-6, sign: float = 1.0, q_chunk: Optional[int] = None):
        _check_2d("X", X)
        _check_2d("Y", Y)
        if Y.shape[0] != X.shape[0] or Y.shape[1] != 1:
            raise ValueError(f"Y must be (n,1) with n={X.shape[0]}, got {tuple(Y.shape)}")
        self._lib = _lib.load()
        self._device = _lib.require_cuda(device)
        self._mean_module = (mean_module if mean_module is not None else ConstantMean()).to(self._device)
        self._cov_module = (cov_module if cov_module is not None else Matern52Kernel()).to(self._device)
        self._warps, (self._kind, self._eps, self._raw_ls, self._raw_var) = split_warps(self._cov_module)
        self._lr, self._epochs = float(lr), int(epochs)
        self._noise_variance, self._sign = float(noise_variance), float(sign)
        self._q_chunk = q_chunk
        p0 = next(iter(list(self._cov_module.parameters()) + list(self._mean_module.parameters())))
        self._mdtype = p0.dtype                        # dtype the torch modules compute in
        self._X = _as_2d_f64(X, self._device)
        self._Y = _as_2d_f64(Y, self._device)          # (n,1)
        self.n, self.d = self._X.shape
        self._ws = {}
        self._inner: Optional[CudaGP] = None
        self.last_values = None

    # ------------------------------------------------------------------ helpers
    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            self._ws[key] = ws
        return ws

    def warp(self, X: torch.Tensor) -> torch.Tensor:
        """kernels.py:65-67 (differentiable; module dtype)."""
        for w in self._warps:
            X = w(X)
        return X

    def _warp_pool(self, Xq: torch.Tensor) -> torch.Tensor:
        """Scoring prologue (no grad): (q,d) device pool -> (q,dw) float64/float32 warped pool."""
        with torch.no_grad():
            out = Xq
            for w in self._warps:
                if hasattr(w, "alpha") and hasattr(w, "beta") and hasattr(w, "eps") and not list(w.children()):
                    a = float(torch.nn.functional.softplus(w.alpha.detach().to(_F64)))
                    b = float(torch.nn.functional.softplus(w.beta.detach().to(_F64)))
                    src = out.contiguous()
                    dst = torch.empty_like(src)
                    with torch.cuda.device(self._device):
                        _lib.check(self._lib.bbo_warp_kumaraswamy(
                            _lib.ptr(src), src.numel(), a, b, float(w.eps), _lib.ptr(dst),
                            int(src.dtype == torch.float32), _lib.stream_ptr(self._device)), "bbo_warp_kumaraswamy")
                    out = dst
                else:
                    was_training = w.training
                    w.eval()
                    out = w(out.to(self._mdtype)).to(out.dtype if out.dtype in (torch.float32, _F64) else _F64)
                    w.train(was_training)
            return out

    def _const_mean(self) -> bool:
        return hasattr(self._mean_module, "constant") and not list(self._mean_module.children())

    # ------------------------------------------------------------------ fit
    def objective(self) -> torch.Tensor:
        """objectives.py:34-45 (the reference's ``GP._objective`` call at gp.py:60) as a differentiable
        (1,1) tensor: ``.backward()`` fills ``.grad`` of every mean / warp / kernel parameter."""
        Xm = self._X.to(self._mdtype)
        Xw = self.warp(Xm)
        resid = self._Y.to(self._mdtype) - self._mean_module(Xm)
        return _Objective.apply(self, Xw, resid.reshape(-1), self._raw_ls, self._raw_var)

    def train(self):
        """gp.py:54-63: Adam(lr) x epochs over mean.parameters() + cov.parameters(), minimising
        ``sign`` x objective (sign=+1 is the reference)."""
        params = list(self._mean_module.parameters()) + list(self._cov_module.parameters())
        opt = torch.optim.Adam(params, lr=self._lr)
        self._mean_module.train()
        self._cov_module.train()
        values = []
        for _ in range(self._epochs):
            opt.zero_grad()
            loss = self.objective()
            values.append(loss.detach().reshape(()))
            (self._sign * loss).sum().backward()
            opt.step()
        self.last_values = torch.stack(values).to(_F64) if values else None
        self._inner = None

    # ------------------------------------------------------------------ posterior
    def _ensure_inner(self) -> CudaGP:
        if self._inner is None:
            self._mean_module.eval()
            self._cov_module.eval()
            with torch.no_grad():
                Xm = self._X.to(self._mdtype)
                Xw = self.warp(Xm).to(_F64)
                base = TemplateKernel({(0, 0.0): "rbf_vmap", (1, 0.0): "matern52_cdist"}.get(
                    (self._kind, self._eps), "matern52_vmap"),
                    ard_dim=(self._raw_ls.numel() if self._raw_ls.dim() == 1 else None)).to(self._device).double()
                base.lengthscale.copy_(self._raw_ls.detach().to(_F64))
                cov = base
                if self._raw_var is not None:
                    cov = ScaleKernel(base).to(self._device).double()
                    cov.variance.copy_(self._raw_var.detach().to(_F64))
                mean = ConstantMean().to(self._device).double()
                if self._const_mean():
                    mean.constant.copy_(self._mean_module.constant.detach().to(_F64))
                    resid = self._Y
                else:
                    resid = self._Y - self._mean_module(Xm).to(_F64).reshape(-1, 1)
            self._inner = CudaGP(Xw, resid, mean, cov, device=self._device, noise_variance=self._noise_variance,
                                 q_chunk=self._q_chunk)
        return self._inner

    def _mean_offset(self, Xq: torch.Tensor) -> Optional[torch.Tensor]:
        if self._const_mean():
            return None
        with torch.no_grad():
            return self._mean_module(Xq.to(device=self._device, dtype=self._mdtype)).to(_F64).reshape(-1)

    def predict_diag(self, query_X: torch.Tensor):
        if query_X.dim() != 2 or query_X.shape[1] != self.d:
            raise ValueError(f"query must be (q,{self.d}), got {tuple(query_X.shape)}")
        inner = self._ensure_inner()
        xq = query_X.detach().to(device=self._device, dtype=_F64)
        mu, var = inner.predict_diag(self._warp_pool(xq))
        off = self._mean_offset(xq)
        return (mu if off is None else mu + off), var

    def predict(self, query_X: torch.Tensor):
        """gp.py:65-83: (q,d) -> mu (q,1), Sigma (q,q); (q,1,d) -> (q,1,1), (q,1,1)."""
        if not isinstance(query_X, torch.Tensor):
            raise TypeError(f"Expected Tensor, got {type(query_X)}")
        out_dtype = query_X.dtype if query_X.dtype.is_floating_point else _F64
        if query_X.dim() == 3 and query_X.shape[1] == 1:
            mu, var = self.predict_diag(query_X.reshape(query_X.shape[0], -1))
            return mu.reshape(-1, 1, 1).to(out_dtype), var.reshape(-1, 1, 1).to(out_dtype)
        if query_X.dim() != 2:
            raise ValueError(f"query must be (q,d) or (q,1,d), got {tuple(query_X.shape)}")
        inner = self._ensure_inner()
        xq = query_X.detach().to(device=self._device, dtype=_F64)
        mu, cov = inner.predict(self._warp_pool(xq))
        off = self._mean_offset(xq)
        if off is not None:
            mu = mu + off.reshape(-1, 1)
        return mu.to(out_dtype), cov.to(out_dtype)

    def score_pool(self, Xq: torch.Tensor, acq, k: int = 1, *, precision: str = "fp64", index_offset: int = 0,
                   return_all: bool = False):
        """Pool scoring + selection.  With a constant mean this is the fused CUDA path on the warped pool;
        with a learned mean the mean offset is added to mu before the acquisition (bbo_acq_eval) and the
        selection is a stable descending sort (ties -> lowest index, trial.py:289-294)."""
        if Xq.dim() != 2 or Xq.shape[1] != self.d:
            raise ValueError(f"pool must be (q,{self.d}), got {tuple(Xq.shape)}")
        inner = self._ensure_inner()
        xq = Xq.detach().to(self._device)
        if xq.dtype not in (torch.float32, _F64):
            xq = xq.to(_F64)
        xw = self._warp_pool(xq)
        if self._const_mean():
            return inner.score_pool(xw, acq, k, precision=precision, index_offset=index_offset,
                                    return_all=return_all)
        from .acquisitions import as_descriptor
        _, _, mu, var, _ = inner.score_pool(xw, acq, 1, precision=precision, return_all=True)
        mu = mu + self._mean_offset(xq)
        sd = var.clamp_min(0).sqrt() if precision == "tf32" else var.sqrt()
        sc = torch.empty_like(mu)
        a = as_descriptor(acq)
        with torch.cuda.device(self._device):
            _lib.check(self._lib.bbo_acq_eval(C.byref(a), _lib.ptr(mu), _lib.ptr(sd), _lib.ptr(sc), mu.numel(), 0,
                                              _lib.stream_ptr(self._device)), "bbo_acq_eval")
        order = torch.sort(torch.nan_to_num(sc, nan=float("-inf")), descending=True, stable=True).indices[:k]
        top_s, top_i = sc[order], order + index_offset
        if return_all:
            return top_s, top_i, mu, var, sc
        return top_s, top_i


def surrogate(X, Y, mean_module=None, cov_module=None, **kw):
    """CudaGP for ConstantMean + plain kernels (fully fused device-side training), CudaWarpedGP when the
    covariance contains a WarpKernel or the mean is not a ConstantMean."""
    cov_has_warp = False
    m = cov_module
    while m is not None and hasattr(m, "base_kernel"):
        cov_has_warp |= hasattr(m, "warp_module")
        m = m.base_kernel
    mean_is_const = mean_module is None or (hasattr(mean_module, "constant") and not list(mean_module.children()))
    if cov_has_warp or not mean_is_const:
        return CudaWarpedGP(X, Y, mean_module, cov_module, **kw)
    return CudaGP(X, Y, mean_module, cov_module, **kw)
