"""CudaGPBatch: many independent small GPs fitted in lock-step (BASELINE config 4:
"256 independent GP fits (n=256, d=10) scored concurrently").

Every kernel of the fit path carries a batch index (blockIdx.z / .y), so B problems cost the same
number of launches as one: per Adam epoch B covariance builds, B blocked Cholesky factorisations,
B TRTRI / LAUUM / gradient reductions and B Adam updates run as single batched launches
(bbo_gp_fit_adam with batch = B).  Problems share the kernel family and (n, d); each has its own
data and hyper-parameters.  The arithmetic per problem is exactly that of ``CudaGP``
(gp/gp.py:54-83, objectives.py:20-45 of the reference).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib
from .gp import CudaGP, _F64, _hyper, _raise_not_pd
from .kernels import IMPLS, ConstantMean, ScaleKernel, TemplateKernel


class CudaGPBatch:
    def __init__(self, X: torch.Tensor, Y: torch.Tensor, impl: str = "matern52_vmap", *, ard: bool = True,
                 scale: bool = True, theta: Optional[torch.Tensor] = None, device="cuda", lr: float = 0.01,
                 epochs: int = 100, noise_variance: float = 1e-6, sign: float = 1.0):
        if X.dim() != 3 or Y.dim() != 3 or Y.shape[:2] != X.shape[:2] or Y.shape[2] != 1:
            raise ValueError(f"expected X (B,n,d) and Y (B,n,1), got {tuple(X.shape)} and {tuple(Y.shape)}")
        if impl not in IMPLS:
            raise ValueError(f"unknown kernel impl {impl!r}")
        self._lib = _lib.load()
        self._device = _lib.require_cuda(device)
        self.impl, self.ard, self.scale = impl, bool(ard), bool(scale)
        self.kind, self.eps = IMPLS[impl]
        self.B, self.n, self.d = X.shape
        self.np = int(self._lib.bbo_padded_n(self.n))
        self._X = X.detach().to(device=self._device, dtype=_F64).contiguous()
        self._Y = Y.detach().to(device=self._device, dtype=_F64).reshape(self.B, self.n).contiguous()
        self.n_ls = self.d if self.ard else 1
        self.P = 1 + self.n_ls + (1 if self.scale else 0)
        if theta is None:   # reference initialisation: constant 0, raw lengthscale/variance ~ N(0,1)
            theta = torch.randn(self.B, self.P, dtype=_F64)
            theta[:, 0] = 0.0
        if tuple(theta.shape) != (self.B, self.P):
            raise ValueError(f"theta must be ({self.B},{self.P}) = [constant, raw_lengthscale.., raw_variance?]")
        self.theta = theta.detach().to(device=self._device, dtype=_F64).contiguous().clone()
        self._lr, self._epochs, self._noise, self._sign = float(lr), int(epochs), float(noise_variance), float(sign)
        self._ws = None
        self._state = None
        self.last_values = None

    def _hyper(self):
        return _hyper(self.kind, self.d, self.eps, self._noise)

    def _workspace(self):
        nbytes = int(self._lib.bbo_gp_fit_workspace_bytes(self.n, self.d, self.B))
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
        return self._ws

    def _hyp(self):
        hyp = torch.empty(self.B, 2 * self.d + 2, dtype=_F64, device=self._device)
        _lib.check(self._lib.bbo_gp_theta_to_hyp(self.d, self.B, int(self.ard), int(self.scale), _lib.ptr(self.theta),
                                                 _lib.ptr(hyp), _lib.stream_ptr(self._device)))
        return hyp

    def value_and_grad(self):
        """(values (B,), grads (B, d+2)) w.r.t. the constrained (l_k, s2, c); see bbo_gp_fit_value_grad."""
        with torch.cuda.device(self._device):
            h, hyp, ws = self._hyper(), self._hyp(), self._workspace()
            value = torch.empty(self.B, dtype=_F64, device=self._device)
            grad = torch.empty(self.B, self.d + 2, dtype=_F64, device=self._device)
            info = torch.zeros(self.B, dtype=torch.int32, device=self._device)
            _lib.check(self._lib.bbo_gp_fit_value_grad(
                C.byref(h), self.n, self.B, _lib.ptr(self._X), _lib.ptr(self._Y), _lib.ptr(hyp), _lib.ptr(value),
                _lib.ptr(grad), _lib.ptr(info), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)),
                "bbo_gp_fit_value_grad")
            bad = torch.nonzero(info)
            if bad.numel():
                _raise_not_pd(int(info[bad[0, 0]]))
        return value, grad

    def train(self):
        """GP.train (gp.py:54-63) for all B problems at once; no host synchronisation inside the loop."""
        with torch.cuda.device(self._device):
            m = torch.zeros_like(self.theta)
            v = torch.zeros_like(self.theta)
            values = torch.empty(max(self._epochs, 1), self.B, dtype=_F64, device=self._device)
            info = torch.zeros(self.B, dtype=torch.int32, device=self._device)
            h, ws = self._hyper(), self._workspace()
            _lib.check(self._lib.bbo_gp_fit_adam(
                C.byref(h), self.n, self.B, int(self.ard), int(self.scale), _lib.ptr(self._X), _lib.ptr(self._Y),
                _lib.ptr(self.theta), _lib.ptr(m), _lib.ptr(v), 0, self._epochs, self._lr, self._sign,
                _lib.ptr(values), _lib.ptr(info), _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)),
                "bbo_gp_fit_adam")
            self.last_values = values[:self._epochs]
            self._state = None
            self.info = info          # per-problem status; non-PD problems keep their last valid parameters
        return self

    def factorize(self):
        if self._state is None:
            with torch.cuda.device(self._device):
                h, hyp, ws = self._hyper(), self._hyp(), self._workspace()
                linv = torch.empty(self.B, self.np, self.np, dtype=_F64, device=self._device)
                alpha = torch.empty(self.B, self.np, dtype=_F64, device=self._device)
                logdet = torch.empty(self.B, dtype=_F64, device=self._device)
                info = torch.zeros(self.B, dtype=torch.int32, device=self._device)
                _lib.check(self._lib.bbo_gp_factorize(
                    C.byref(h), self.n, self.B, _lib.ptr(self._X), _lib.ptr(self._Y), _lib.ptr(hyp), _lib.ptr(linv),
                    _lib.ptr(alpha), None, _lib.ptr(logdet), _lib.ptr(info), _lib.ptr(ws), ws.numel(),
                    _lib.stream_ptr(self._device)), "bbo_gp_factorize")
                self._state = {"hyp": hyp, "linv": linv, "alpha": alpha, "logdet": logdet, "info": info}
        return self._state

    def problem(self, b: int, q_chunk: Optional[int] = None) -> CudaGP:
        """A ``CudaGP`` view of problem b sharing the batched factorisation (no refit)."""
        s = self.factorize()
        th = self.theta[b]
        base = TemplateKernel(self.impl, ard_dim=self.d if self.ard else None).double()
        with torch.no_grad():
            base.lengthscale.copy_(th[1:1 + self.n_ls].reshape(base.lengthscale.shape))
        cov = base
        if self.scale:
            cov = ScaleKernel(base).double()
            with torch.no_grad():
                cov.variance.copy_(th[1 + self.n_ls])
        mean = ConstantMean().double()
        with torch.no_grad():
            mean.constant.copy_(th[0])
        gp = CudaGP(self._X[b], self._Y[b].reshape(-1, 1), mean, cov, device=self._device,
                    noise_variance=self._noise, lr=self._lr, epochs=self._epochs, sign=self._sign, q_chunk=q_chunk)
        gp.install_state(s["hyp"][b], s["linv"][b], s["alpha"][b])
        return gp

    def _score_pools_batched(self, Xq: torch.Tensor, acq_factory, k: int):
        """All problems in one set of launches (``bbo_gp_score_pool_batched``, blockIdx.z = problem)."""
        from .acquisitions import as_descriptor
        s = self._state
        q = Xq.shape[1]
        with torch.cuda.device(self._device):
            xq = Xq.detach().to(device=self._device, dtype=_F64).contiguous()
            arr = (_lib.Acq * self.B)(*[as_descriptor(acq_factory(b)) for b in range(self.B)])
            acq_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self._device)
            out_s = torch.empty(self.B, k, dtype=_F64, device=self._device)
            out_i = torch.empty(self.B, k, dtype=torch.int64, device=self._device)
            nbytes = int(self._lib.bbo_gp_score_batched_workspace_bytes(self.n, self.d, self.B, q))
            ws = torch.empty(nbytes, dtype=torch.uint8, device=self._device)
            h = self._hyper()
            _lib.check(self._lib.bbo_gp_score_pool_batched(
                C.byref(h), self.n, self.B, _lib.ptr(self._X), _lib.ptr(s["hyp"]), _lib.ptr(s["linv"]),
                _lib.ptr(s["alpha"]), _lib.ptr(acq_dev), _lib.ptr(xq), q, k, _lib.ptr(out_s), _lib.ptr(out_i),
                _lib.ptr(ws), ws.numel(), _lib.stream_ptr(self._device)), "bbo_gp_score_pool_batched")
        return out_s, out_i

    def score_pools(self, Xq: torch.Tensor, acq_factory, k: int = 1, *, precision: str = "fp64", n_streams: int = 8):
        """Score one pool per problem concurrently: Xq (B,q,d); acq_factory(b) -> acquisition.
        FP64 mode: one batched library call for all problems (``bbo_gp_score_pool_batched``).  TF32 mode (or
        pools too large for the single-merge selection): problems are issued round-robin on `n_streams` CUDA
        streams so that the per-problem kernels overlap.  Returns (scores (B,k), indices (B,k))."""
        if Xq.dim() != 3 or Xq.shape[0] != self.B or Xq.shape[2] != self.d:
            raise ValueError(f"pools must be ({self.B},q,{self.d})")
        self.factorize()
        q = Xq.shape[1]
        if precision == "fp64" and q >= 1 and (q // 2048 + 2) * k <= 2048:
            return self._score_pools_batched(Xq, acq_factory, k)
        gps: List[CudaGP] = [self.problem(b, q_chunk=max(128, (q + 127) // 128 * 128)) for b in range(self.B)]
        streams = [torch.cuda.Stream(device=self._device) for _ in range(max(1, n_streams))]
        cur = torch.cuda.current_stream(self._device)
        out_s = torch.empty(self.B, k, dtype=_F64, device=self._device)
        out_i = torch.empty(self.B, k, dtype=torch.int64, device=self._device)
        xq = Xq.detach().to(self._device)
        for st in streams:
            st.wait_stream(cur)
        for b, gp in enumerate(gps):
            st = streams[b % len(streams)]
            with torch.cuda.stream(st):
                s, i = gp.score_pool(xq[b], acq_factory(b), k, precision=precision)
                out_s[b].copy_(s)
                out_i[b].copy_(i)
        for st in streams:
            cur.wait_stream(st)
        self._keep = gps      # keep workspaces alive until the streams have drained
        return out_s, out_i
