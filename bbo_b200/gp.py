"""CudaGP: the reference's native GP surrogate (bbo/algorithms/surrogates/gp/gp.py:30-83)
with every numeric step executed by libbbo_b200.so on a B200.

Same constructor shape, same ``train()`` / ``predict(X)`` contract:
  * ``train``   gp.py:54-63   Adam(lr) x epochs on [mean.parameters(), cov.parameters()],
                              *minimising the returned objective value*, which in the reference is
                              +log p(y|X).  ``sign=+1`` (default) reproduces that;
                              ``sign=-1`` minimises the true negative log-likelihood.
  * ``predict`` gp.py:65-83   (q,d) -> mu (q,1), Sigma (q,q);  (q,1,d) -> (q,1,1), (q,1,1)
                              (the reference's 3-D path is broken; the shape convention
                              implemented is the one BODesigner.score_fn relies on, bo.py:112-114).
Extras used by the pool optimiser: ``predict_diag`` and ``score_pool``.

All arithmetic is float64 on the device (>= the reference's precision for any input dtype),
except the optional TF32 scoring mode of ``score_pool``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import os

import torch
import torch.nn as nn

from . import _lib
from .abstractions import Surrogate, adopt_reference_abcs
from .kernels import ConstantMean, Matern52Kernel, describe_cov, describe_mean

_F64 = torch.float64


def _as_2d_f64(x: torch.Tensor, device: torch.device) -> torch.Tensor:
    return x.detach().to(device=device, dtype=_F64).contiguous()


def _check_2d(name: str, value):
    # bbo/utils/attrs_utils.py:61-65
    if not isinstance(value, torch.Tensor):
        raise TypeError(f"Expected Tensor, got {type(value)}")
    if value.dim() != 2:
        raise ValueError(f"Expected 2D Tensor, got {value.dim()}D Tensor")


_KS_BUDGET_MB = int(os.environ.get("BBO_TF32_KS_BUDGET_MB", "768"))   # K* bytes (binary16) per tensor-core chunk


def default_q_chunk(n: int, budget_bytes: int = 512 << 20) -> int:
    np_ = (n + _lib.ROW_ALIGN - 1) // _lib.ROW_ALIGN * _lib.ROW_ALIGN
    q = budget_bytes // (np_ * 8)
    q = max(1024, min(65536, q))
    return q // 128 * 128


class _Theta:
    """Packs module parameters into the Adam-ordered raw vector (gp.py:55) and back."""

    def __init__(self, mean_module: nn.Module, cov_module: nn.Module):
        self.kind, self.eps, self.raw_ls, self.raw_var = describe_cov(cov_module)
        self.const = describe_mean(mean_module)
        self.ard = self.raw_ls.dim() == 1
        self.n_ls = self.raw_ls.numel() if self.ard else 1
        self.has_scale = self.raw_var is not None
        self.P = 1 + self.n_ls + (1 if self.has_scale else 0)

    def pack(self, device) -> torch.Tensor:
        parts = [self.const.detach().reshape(1), self.raw_ls.detach().reshape(-1)]
        if self.has_scale:
            parts.append(self.raw_var.detach().reshape(1))
        return torch.cat([p.to(device=device, dtype=_F64) for p in parts]).contiguous()

    def unpack(self, theta: torch.Tensor) -> None:
        with torch.no_grad():
            self.const.copy_(theta[0].to(self.const.dtype).reshape(self.const.shape))
            self.raw_ls.copy_(theta[1:1 + self.n_ls].to(self.raw_ls.dtype).reshape(self.raw_ls.shape))
            if self.has_scale:
                self.raw_var.copy_(theta[1 + self.n_ls].to(self.raw_var.dtype).reshape(self.raw_var.shape))


def _hyper(kind: int, d: int, eps: float, noise: float) -> _lib.GpHyper:
    # The reference adds ``torch.eye(n) * noise_variance`` (objectives.py:28, gp.py:82): a float32 tensor, so
    # the noise that reaches its float64 K is float32(noise).  Reproduced here (|rel. change| <= 6e-8).
    noise32 = float(torch.tensor(noise, dtype=torch.float32))
    return _lib.GpHyper(kind=kind, d=d, pairwise_eps=eps, noise=noise32)


def _raise_not_pd(info: int):
    raise torch.linalg.LinAlgError(
        f"linalg.cholesky: The factorization could not be completed because the input is not "
        f"positive-definite (the leading minor of order {info} is not positive-definite).")


def cov_matrix(cov_module: nn.Module, X1: torch.Tensor, X2: torch.Tensor) -> torch.Tensor:
    """KernelProtocol (kernel_impl.py:22-24) on the device: (q,d) x (n,d) -> (q,n);
    3-D operands are batched like kernel_impl.py:106-128."""
    lib = _lib.load()
    dev = _lib.require_cuda(X1.device if X1.is_cuda else None)
    if X1.dim() == 3 or X2.dim() == 3:
        b = X1.shape[0] if X1.dim() == 3 else X2.shape[0]
        A = X1 if X1.dim() == 3 else X1.expand(b, -1, -1)
        B = X2 if X2.dim() == 3 else X2.expand(b, -1, -1)
        return torch.stack([cov_matrix(cov_module, A[i], B[i]) for i in range(b)])
    if X1.dim() != 2 or X2.dim() != 2:
        raise ValueError("X1 and X2 must be 2d or 3d tensors. Got {} and {}.".format(X1.shape, X2.shape))
    th = _Theta(ConstantMean(), cov_module)
    d = X1.shape[1]
    theta = th.pack(dev)
    hyp = torch.empty(2 * d + 2, dtype=_F64, device=dev)
    st = _lib.stream_ptr(dev)
    _lib.check(lib.bbo_gp_theta_to_hyp(d, 1, int(th.ard), int(th.has_scale), _lib.ptr(theta), _lib.ptr(hyp), st))
    a, b = _as_2d_f64(X1, dev), _as_2d_f64(X2, dev)
    out = torch.empty(a.shape[0], b.shape[0], dtype=_F64, device=dev)
    h = _hyper(th.kind, d, th.eps, 0.0)
    _lib.check(lib.bbo_cov_matrix(C.byref(h), _lib.ptr(hyp), _lib.ptr(a), a.shape[0], _lib.ptr(b), b.shape[0],
                                  _lib.ptr(out), b.shape[0], st), "bbo_cov_matrix")
    return out.to(X1.dtype) if X1.dtype.is_floating_point else out


class CudaGP(Surrogate):
    def __init__(self, X: torch.Tensor, Y: torch.Tensor, mean_module: Optional[nn.Module] = None,
                 cov_module: Optional[nn.Module] = None, *, device="cuda", lr: float = 0.01,
                 epochs: int = 100, noise_variance: float = 1e-6, sign: float = 1.0,
                 q_chunk: Optional[int] = None, out_dtype: Optional[torch.dtype] = None):
        """``out_dtype``: dtype of what ``predict`` returns.  None (default) follows the query's dtype, as the
        reference does; ``torch.float64`` keeps the library's float64 results even for the float32 features the
        reference's BODesigner sends, so that the acquisition is evaluated in float64 too."""
        _check_2d("X", X)
        _check_2d("Y", Y)
        if Y.shape[0] != X.shape[0] or Y.shape[1] != 1:
            raise ValueError(f"Y must be (n,1) with n={X.shape[0]}, got {tuple(Y.shape)}")
        self._lib = _lib.load()
        adopt_reference_abcs(type(self), "Surrogate")
        self._device = _lib.require_cuda(device)
        self._mean_module = (mean_module if mean_module is not None else ConstantMean()).to(self._device)
        self._cov_module = (cov_module if cov_module is not None else Matern52Kernel()).to(self._device)
        self._lr = float(lr)
        self._epochs = int(epochs)
        self._noise_variance = float(noise_variance)
        self._sign = float(sign)
        self._X = _as_2d_f64(X, self._device)
        self._Y = _as_2d_f64(Y, self._device).reshape(-1).contiguous()
        self.n, self.d = self._X.shape
        self.np = int(self._lib.bbo_padded_n(self.n))
        self._theta = _Theta(self._mean_module, self._cov_module)
        if self._theta.ard and self._theta.n_ls != self.d:
            raise ValueError(f"ARD lengthscale has {self._theta.n_ls} dims, X has {self.d}")
        # chunks are multiples of 128 candidates, at least 128 (the library rounds the same way)
        self._q_chunk = max(128, (int(q_chunk) + 127) // 128 * 128) if q_chunk else default_q_chunk(self.n)
        self._q_chunk_auto = not q_chunk
        self._out_dtype = out_dtype
        self._sm_count = torch.cuda.get_device_properties(self._device).multi_processor_count
        self._last_refine = None
        self._state = None           # (GpState, keep-alive tensors)
        self._ws = {}                # workspace cache
        self.last_values = None      # objective value seen at each epoch of the last train()
        # Adam state survives across train() calls only if the caller asks (warm start)
        self._adam = None

    # ------------------------------------------------------------------ helpers
    def _hyper(self) -> _lib.GpHyper:
        return _hyper(self._theta.kind, self.d, self._theta.eps, self._noise_variance)

    def _workspace(self, key: str, nbytes: int) -> torch.Tensor:
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            self._ws[key] = ws
        return ws

    def _fit_ws(self) -> torch.Tensor:
        return self._workspace("fit", int(self._lib.bbo_gp_fit_workspace_bytes(self.n, self.d, 1)))

    def _hyp_from_modules(self) -> torch.Tensor:
        theta = self._theta.pack(self._device)
        hyp = torch.empty(2 * self.d + 2, dtype=_F64, device=self._device)
        _lib.check(self._lib.bbo_gp_theta_to_hyp(self.d, 1, int(self._theta.ard), int(self._theta.has_scale),
                                                 _lib.ptr(theta), _lib.ptr(hyp), _lib.stream_ptr(self._device)))
        return hyp

    # ------------------------------------------------------------------ fit
    def value_and_grad(self):
        """objectives.py:34-45 value (= +log p(y|X), F4) and its gradient w.r.t. the RAW
        parameters, as autograd would return them: (value, g_raw_lengthscale, g_raw_variance|None,
        g_constant) -- float64 device tensors."""
        with torch.cuda.device(self._device):
            h = self._hyper()
            ws = self._fit_ws()
            # the library replays the whole iteration as a CUDA graph while it is called with the same buffers, so
            # the hyper-parameters, outputs and the info word live in buffers kept with the workspace; the caller
            # gets copies
            bufs = self._ws.get("vg")
            if bufs is None or bufs[0].numel() != 2 * self.d + 2:
                bufs = (torch.empty(2 * self.d + 2, dtype=_F64, device=self._device),
                        torch.empty(1, dtype=_F64, device=self._device),
                        torch.empty(self.d + 2, dtype=_F64, device=self._device),
                        torch.zeros(1, dtype=torch.int32, device=self._device))
                self._ws["vg"] = bufs
            hyp, value_b, grad_b, info = bufs
            hyp.copy_(self._hyp_from_modules())
            info.zero_()
            _lib.check(self._lib.bbo_gp_fit_value_grad(
                C.byref(h), self.n, 1, _lib.ptr(self._X), _lib.ptr(self._Y), _lib.ptr(hyp), _lib.ptr(value_b),
                _lib.ptr(grad_b), _lib.ptr(info), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)),
                "bbo_gp_fit_value_grad")
            i = int(info.item())
            if i != 0:
                _raise_not_pd(i)
            value, grad = value_b.clone(), grad_b.clone()
        th = self._theta
        raw_ls = th.raw_ls.detach().to(device=self._device, dtype=_F64)
        sg = torch.where(raw_ls > 20.0, torch.ones_like(raw_ls), torch.sigmoid(raw_ls))
        g_ls = grad[:self.d] * sg if th.ard else (grad[:self.d].sum() * sg)
        g_var = None
        if th.has_scale:
            rv = th.raw_var.detach().to(device=self._device, dtype=_F64)
            g_var = grad[self.d] * torch.where(rv > 20.0, torch.ones_like(rv), torch.sigmoid(rv))
        return value[0], g_ls, g_var, grad[self.d + 1]

    def train(self, X=None, Y=None, *, warm_start: bool = False):
        """GP.train (gp.py:54-63).  The whole epochs x (factorise, gradient, Adam) loop is enqueued on
        the current stream without host synchronisation; parameters are written back to the modules.
        ``X`` / ``Y`` exist only because the ``Surrogate`` ABC declares them (abstractions.py:78-81); like the
        reference's ``GP.train()`` the data given to the constructor is used, so passing them is an error."""
        if X is not None or Y is not None:
            raise TypeError("CudaGP.train() uses the data given to the constructor (as GP.train does, gp.py:54); "
                            "build a new CudaGP or use append() for new observations")
        with torch.cuda.device(self._device):
            th = self._theta
            theta = th.pack(self._device)
            if warm_start and self._adam is not None:
                m, v, step0 = self._adam
            else:
                m = torch.zeros(th.P, dtype=_F64, device=self._device)
                v = torch.zeros(th.P, dtype=_F64, device=self._device)
                step0 = 0
            values = torch.empty(max(self._epochs, 1), dtype=_F64, device=self._device)
            info = torch.zeros(1, dtype=torch.int32, device=self._device)
            ws = self._fit_ws()
            h = self._hyper()
            _lib.check(self._lib.bbo_gp_fit_adam(
                C.byref(h), self.n, 1, int(th.ard), int(th.has_scale), _lib.ptr(self._X), _lib.ptr(self._Y),
                _lib.ptr(theta), _lib.ptr(m), _lib.ptr(v), step0, self._epochs, self._lr, self._sign,
                _lib.ptr(values), _lib.ptr(info), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)),
                "bbo_gp_fit_adam")
            self._adam = (m, v, step0 + self._epochs)
            th.unpack(theta)
            self.last_values = values[:self._epochs]
            self._state = None   # deliberate: the reference keeps a stale cache here (gp.py:44-45,66)
            i = int(info.item())
            if i != 0:
                _raise_not_pd(i)

    # ------------------------------------------------------------------ posterior state
    def _ensure_state(self, want_tf32: bool = False):
        if self._state is None:
            with torch.cuda.device(self._device):
                h = self._hyper()
                hyp = self._hyp_from_modules()
                linv = torch.empty(self.np, self.np, dtype=_F64, device=self._device)
                alpha = torch.empty(self.np, dtype=_F64, device=self._device)
                logdet = torch.empty(1, dtype=_F64, device=self._device)
                info = torch.zeros(1, dtype=torch.int32, device=self._device)
                ws = self._fit_ws()
                _lib.check(self._lib.bbo_gp_factorize(
                    C.byref(h), self.n, 1, _lib.ptr(self._X), _lib.ptr(self._Y), _lib.ptr(hyp), _lib.ptr(linv),
                    _lib.ptr(alpha), None, _lib.ptr(logdet), _lib.ptr(info), _lib.ptr(ws), ws.numel(),
                    _lib.stream_ptr(self._device)), "bbo_gp_factorize")
                i = int(info.item())
                if i != 0:
                    _raise_not_pd(i)
                st = _lib.GpState(h=h, n=self.n, X=self._X.data_ptr(), hyp=hyp.data_ptr(),
                                  linv=linv.data_ptr(), alpha=alpha.data_ptr(), linv_tf32=None)
                self._state = {"st": st, "hyp": hyp, "linv": linv, "alpha": alpha, "logdet": logdet,
                               "linv_tf32": None}
        if want_tf32 and self._state["linv_tf32"] is None:
            with torch.cuda.device(self._device):
                lt = torch.empty(int(self._lib.bbo_tf32_rows(self.n)), self.np, dtype=torch.float32,
                                 device=self._device)
                _lib.check(self._lib.bbo_gp_make_tf32(self.n, _lib.ptr(self._state["linv"]), _lib.ptr(lt),
                                                      _lib.stream_ptr(self._device)), "bbo_gp_make_tf32")
                self._state["linv_tf32"] = lt
                self._state["st"].linv_tf32 = lt.data_ptr()
        return self._state

    @property
    def state(self):
        return self._ensure_state()

    def install_state(self, hyp: torch.Tensor, linv: torch.Tensor, alpha: torch.Tensor) -> None:
        """Adopt a fitted state computed elsewhere (e.g. broadcast from the fitting GPU)."""
        if linv.shape != (self.np, self.np) or alpha.shape != (self.np,) or hyp.shape != (2 * self.d + 2,):
            raise ValueError("state tensors have the wrong shape")
        st = _lib.GpState(h=self._hyper(), n=self.n, X=self._X.data_ptr(), hyp=hyp.data_ptr(),
                          linv=linv.data_ptr(), alpha=alpha.data_ptr(), linv_tf32=None)
        self._state = {"st": st, "hyp": hyp, "linv": linv, "alpha": alpha, "logdet": None, "linv_tf32": None}

    # ------------------------------------------------------------------ incremental update
    def append(self, X_new: torch.Tensor, Y_new: torch.Tensor, *, Y_all: Optional[torch.Tensor] = None) -> None:
        """Add observations WITHOUT refitting the hyper-parameters: the factorised state
        (L^-1, alpha) is extended by ``bbo_gp_append`` in O(m n^2) instead of the O(n^3) refactorisation
        (and instead of the reference's full refit, bo.py:101-106).  The result is the state the full
        factorisation of the enlarged data set would give with the current hyper-parameters.
        ``Y_all`` (n+m,1), if given, replaces ALL labels (BODesigner re-standardises the labels at every
        suggestion, bo.py:126-128; L^-1 does not depend on them and alpha is recomputed from scratch).
        Opt-in: call ``train()`` whenever the hyper-parameters should follow the data again."""
        _check_2d("X_new", X_new)
        _check_2d("Y_new", Y_new)
        if X_new.shape[1] != self.d or Y_new.shape != (X_new.shape[0], 1):
            raise ValueError(f"expected X_new (m,{self.d}) and Y_new (m,1), got {tuple(X_new.shape)}, "
                             f"{tuple(Y_new.shape)}")
        xn = _as_2d_f64(X_new, self._device)
        yn = _as_2d_f64(Y_new, self._device).reshape(-1)
        if Y_all is not None:
            _check_2d("Y_all", Y_all)
            if Y_all.shape != (self.n + xn.shape[0], 1):
                raise ValueError(f"Y_all must be ({self.n + xn.shape[0]},1), got {tuple(Y_all.shape)}")
            y_all = _as_2d_f64(Y_all, self._device).reshape(-1)
            self._Y = y_all[:self.n].contiguous()
            yn = y_all[self.n:].contiguous()
        if xn.shape[0] == 0:
            if Y_all is not None:
                self._state = None       # alpha depends on the labels: re-factorise lazily with the new ones
            return
        s = self._ensure_state()
        with torch.cuda.device(self._device):
            for a in range(0, xn.shape[0], _lib.APPEND_MAX):
                xb, yb = xn[a:a + _lib.APPEND_MAX], yn[a:a + _lib.APPEND_MAX]
                m, n_old = xb.shape[0], self.n
                X_all = torch.cat([self._X, xb]).contiguous()
                Y_all = torch.cat([self._Y, yb]).contiguous()
                np_new = int(self._lib.bbo_padded_n(n_old + m))
                linv, alpha = s["linv"], s["alpha"]
                if np_new != self.np:      # re-lay out: old block top-left, unit diagonal on the new padding
                    big = torch.zeros(np_new, np_new, dtype=_F64, device=self._device)
                    big[:self.np, :self.np] = linv
                    big.diagonal()[self.np:] = 1.0
                    linv = big
                    alpha = torch.zeros(np_new, dtype=_F64, device=self._device)
                info = torch.zeros(1, dtype=torch.int32, device=self._device)
                ws = self._workspace("append", int(self._lib.bbo_gp_append_workspace_bytes(n_old, m)))
                h = self._hyper()
                _lib.check(self._lib.bbo_gp_append(
                    C.byref(h), n_old, m, _lib.ptr(X_all), _lib.ptr(Y_all), _lib.ptr(s["hyp"]), _lib.ptr(linv),
                    _lib.ptr(alpha), _lib.ptr(s["logdet"]), _lib.ptr(info), _lib.ptr(ws), ws.numel(),
                    _lib.stream_ptr(self._device)), "bbo_gp_append")
                i = int(info.item())
                if i != 0:
                    self._state = None
                    _raise_not_pd(i)
                self._X, self._Y = X_all, Y_all
                self.n, self.np = n_old + m, np_new
                st = _lib.GpState(h=h, n=self.n, X=self._X.data_ptr(), hyp=s["hyp"].data_ptr(),
                                  linv=linv.data_ptr(), alpha=alpha.data_ptr(), linv_tf32=None)
                s = {"st": st, "hyp": s["hyp"], "linv": linv, "alpha": alpha, "logdet": s["logdet"],
                     "linv_tf32": None}
                self._state = s
            if self._q_chunk_auto:
                self._q_chunk = default_q_chunk(self.n)
            self._ws.pop("fit", None)      # sized for the old n

    # ------------------------------------------------------------------ predict
    def predict_diag(self, query_X: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Per-candidate posterior mean and variance, float64 (q,), (q,)."""
        if query_X.dim() != 2 or query_X.shape[1] != self.d:
            raise ValueError(f"query must be (q,{self.d}), got {tuple(query_X.shape)}")
        s = self._ensure_state()
        with torch.cuda.device(self._device):
            xq = _as_2d_f64(query_X, self._device)
            q = xq.shape[0]
            mu = torch.empty(q, dtype=_F64, device=self._device)
            var = torch.empty(q, dtype=_F64, device=self._device)
            qc = min(self._q_chunk, max(128, (q + 127) // 128 * 128))
            ws = self._workspace("predict", int(self._lib.bbo_gp_predict_workspace_bytes(self.n, self.d, qc, 0)))
            _lib.check(self._lib.bbo_gp_predict(C.byref(s["st"]), _lib.ptr(xq), q, _lib.ptr(mu), _lib.ptr(var),
                                                None, _lib.ptr(ws), ws.numel(), qc,
                                                _lib.stream_ptr(self._device)), "bbo_gp_predict")
        return mu, var

    def _predict_full(self, xq: torch.Tensor):
        s = self._ensure_state()
        with torch.cuda.device(self._device):
            q = xq.shape[0]
            mu = torch.empty(q, dtype=_F64, device=self._device)
            cov = torch.empty(q, q, dtype=_F64, device=self._device)
            qc = max(128, (q + 127) // 128 * 128)
            ws = self._workspace("predict_full",
                                 int(self._lib.bbo_gp_predict_workspace_bytes(self.n, self.d, qc, 1)))
            _lib.check(self._lib.bbo_gp_predict(C.byref(s["st"]), _lib.ptr(xq), q, _lib.ptr(mu), None,
                                                _lib.ptr(cov), _lib.ptr(ws), ws.numel(), qc,
                                                _lib.stream_ptr(self._device)), "bbo_gp_predict(full)")
        return mu, cov

    def predict(self, query_X: torch.Tensor):
        """GP.predict (gp.py:65-83)."""
        if not isinstance(query_X, torch.Tensor):
            raise TypeError(f"Expected Tensor, got {type(query_X)}")
        out_dtype = query_X.dtype if query_X.dtype.is_floating_point else _F64
        if self._out_dtype is not None:
            out_dtype = self._out_dtype
        if query_X.dim() == 2:
            mu, cov = self._predict_full(_as_2d_f64(query_X, self._device))
            return mu.reshape(-1, 1).to(out_dtype), cov.to(out_dtype)
        if query_X.dim() == 3:
            b, m, d = query_X.shape
            if m == 1:   # what BODesigner.score_fn sends (bo.py:112): one candidate per batch entry
                mu, var = self.predict_diag(query_X.reshape(b, d))
                return mu.reshape(b, 1, 1).to(out_dtype), var.reshape(b, 1, 1).to(out_dtype)
            mus, covs = [], []
            for i in range(b):
                mu, cov = self._predict_full(_as_2d_f64(query_X[i], self._device))
                mus.append(mu.reshape(m, 1))
                covs.append(cov)
            return torch.stack(mus).to(out_dtype), torch.stack(covs).to(out_dtype)
        raise ValueError(f"query must be 2-D or 3-D, got {query_X.dim()}-D")

    # ------------------------------------------------------------------ pool scoring
    def score_pool(self, Xq: torch.Tensor, acq, k: int = 1, *, precision: str = "fp64", index_offset: int = 0,
                   return_all: bool = False, running=None, host_pool: bool = False):
        """Score a candidate pool and select the k best (descending, ties -> lowest index).

        Xq: (q,d) float64/float32; a CUDA tensor, or (``host_pool=True``) a CPU tensor that is streamed
        to the device in chunks overlapped with the kernels.  ``acq`` is an EI/UCB/PI from
        ``bbo_b200.acquisitions``.  Returns (scores[k] f64, indices[k] i64[, mu, var, score]).
        ``running`` = (scores, indices) from a previous call to merge with (pool sharded in time).
        ``precision``: "fp64" | "tf32" | "tf32+refine" (TF32 short-list, float64 re-score and selection) | "auto"
        (tf32+refine when its evidence certifies the float64 selection, float64 otherwise; synchronises)."""
        from .acquisitions import as_descriptor
        if Xq.dim() != 2 or Xq.shape[1] != self.d:
            raise ValueError(f"pool must be (q,{self.d}), got {tuple(Xq.shape)}")
        if not (1 <= k <= _lib.TOPK_MAX):
            raise ValueError(f"k must be in [1,{_lib.TOPK_MAX}]")
        if precision == "auto":
            # tensor-core pass + float64 re-score, kept only if the evidence certifies that the selection equals the
            # float64 mode's (refine_certificate); otherwise the pool is scored again in float64
            if return_all or k > _lib.REFINE_KMAX:
                precision = "fp64"
            else:
                out = self.score_pool(Xq, acq, k, precision="tf32+refine", index_offset=index_offset,
                                      running=None if running is None else (running[0].clone(), running[1].clone()),
                                      host_pool=host_pool)
                if self.refine_certificate()["certified"]:
                    self.last_auto_path = "tf32+refine"
                    if running is not None:
                        running[0].copy_(out[0])
                        running[1].copy_(out[1])
                    return out
                precision = "fp64"
            self.last_auto_path = "fp64"
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_lib.PRECISIONS) + ['auto']}")
        prec = _lib.PRECISIONS[precision]
        refine = prec == _lib.PREC_TF32_REFINE
        if refine and k > _lib.REFINE_KMAX:
            raise ValueError(f"precision='tf32+refine' supports k <= {_lib.REFINE_KMAX}")
        if host_pool and (running is not None or return_all):
            raise ValueError("host_pool=True streams the pool from host memory and returns only the selection: "
                             "`running` and `return_all` are not supported with it")
        s = self._ensure_state(want_tf32=prec != _lib.PREC_FP64)
        a = as_descriptor(acq)
        is_f32 = Xq.dtype == torch.float32
        if not is_f32 and Xq.dtype != _F64:
            Xq = Xq.to(_F64)
        q = Xq.shape[0]
        with torch.cuda.device(self._device):
            qc = min(self._q_chunk, max(128, (q + 127) // 128 * 128))
            wave = self._sm_count * 128      # one 128-candidate tile per SM and per wave
            if prec != _lib.PREC_FP64 and q >= wave and self._q_chunk_auto:
                qc = min(q // wave, max(1, (_KS_BUDGET_MB << 20) // (wave * self.np * 2))) * wave
            self._score_shape = (qc, prec)
            ws = self._workspace("score%d" % prec,
                                 int(self._lib.bbo_gp_score_workspace_bytes(self.n, self.d, qc, prec)))
            if running is not None:
                top_s, top_i = running
            else:
                top_s = torch.empty(k, dtype=_F64, device=self._device)
                top_i = torch.empty(k, dtype=torch.int64, device=self._device)
            st = _lib.stream_ptr(self._device)
            if host_pool:
                if Xq.is_cuda:
                    raise ValueError("host_pool=True expects a CPU tensor")
                xq = Xq.contiguous()
                qcr = max(128, (qc + 127) // 128 * 128)      # the library's rounding (bbo_b200.h)
                staging = self._workspace("staging%d" % int(is_f32),
                                          2 * _lib.HOST_STAGE_CHUNKS * qcr * self.d * (4 if is_f32 else 8))
                hs = torch.empty(k, dtype=_F64)
                hi = torch.empty(k, dtype=torch.int64)
                _lib.check(self._lib.bbo_gp_score_pool_host(
                    C.byref(s["st"]), C.byref(a), prec, C.c_void_p(xq.data_ptr()), int(is_f32), q, index_offset,
                    k, C.c_void_p(hs.data_ptr()), C.c_void_p(hi.data_ptr()), _lib.ptr(staging), _lib.ptr(top_s),
                    _lib.ptr(top_i), _lib.ptr(ws), ws.numel(), qc, st), "bbo_gp_score_pool_host")
                return hs, hi
            xq = Xq.detach().to(self._device).contiguous()
            mu = var = sc = None
            if return_all:
                mu = torch.empty(q, dtype=_F64, device=self._device)
                var = torch.empty(q, dtype=_F64, device=self._device)
                sc = torch.empty(q, dtype=_F64, device=self._device)
            nan_count = torch.zeros(1, dtype=torch.int32, device=self._device)
            _lib.check(self._lib.bbo_gp_score_pool(
                C.byref(s["st"]), C.byref(a), prec, _lib.ptr(xq), int(is_f32), q, index_offset, k,
                _lib.ptr(top_s), _lib.ptr(top_i), _lib.ptr(mu), _lib.ptr(var), _lib.ptr(sc), _lib.ptr(nan_count),
                _lib.ptr(ws), ws.numel(), qc, int(running is not None), st), "bbo_gp_score_pool")
            self.last_nan_count = nan_count
        if return_all:
            return top_s, top_i, mu, var, sc
        return top_s, top_i

    def refine_certificate(self, factor: float = 4.0) -> dict:
        """Evidence of the last ``precision="tf32+refine"`` call (``bbo_gp_refine_info``; synchronises):
        ``kth_score`` the k-th best float64 score returned, ``unscored_bound`` an upper bound of the TF32 score of
        every candidate that was not re-scored, ``tf32_error`` the largest |TF32 - float64| score difference seen on
        the ``at_the_cut`` short-listed candidates that did not make the top k (for EI / PI the error scales with
        the score: the top of the list says nothing about the cut; ``tf32_error_whole_list`` is the maximum over
        all of them), ``shortlist`` its size, and ``certified`` = the margin kth_score - unscored_bound exceeds
        ``factor`` x tf32_error (or nothing was left unscored): no unscored candidate can belong to the float64
        top-k, i.e. the selection equals the ``precision="fp64"`` selection."""
        qc, prec = getattr(self, "_score_shape", (None, None))
        if prec != _lib.PREC_TF32_REFINE:
            raise RuntimeError("no tf32+refine call has been made")
        ws = self._ws["score%d" % prec]
        info = (C.c_double * 6)()
        with torch.cuda.device(self._device):
            _lib.check(self._lib.bbo_gp_refine_info(_lib.ptr(ws), ws.numel(), self.n, self.d, qc, info,
                                                    _lib.stream_ptr(self._device)), "bbo_gp_refine_info")
        kth, bound, err, cnt, err_all, ncut = (float(v) for v in info)
        margin = kth - bound
        return {"kth_score": kth, "unscored_bound": bound, "tf32_error": err, "tf32_error_whole_list": err_all,
                "at_the_cut": int(ncut), "shortlist": int(cnt), "margin": margin,
                "certified": bool(bound == float("-inf") or margin > factor * err)}
