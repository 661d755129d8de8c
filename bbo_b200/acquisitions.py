"""Acquisition functions (bbo/algorithms/designers/bo/acquisitions.py:7-31) on the device.

``EI`` and ``UCB`` keep the reference's constructor arguments and ``__call__(mu, stddev)``
contract; ``PI`` is new (absent from the reference) and shares EI's Z.
Called on tensors they run the CUDA epilogue kernel elementwise; handed to
``CudaGP.score_pool`` / ``PoolOptimizer`` they are fused into the scoring pass.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .abstractions import AcquisitionFunction, adopt_reference_abcs


class _DeviceAcq(AcquisitionFunction):
    kind = -1

    def descriptor(self) -> _lib.Acq:
        raise NotImplementedError

    def __call__(self, mu: torch.Tensor, stddev: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        dev = _lib.require_cuda(mu.device if mu.is_cuda else None)
        if self.kind != _lib.ACQ_UCB and bool(torch.isnan(stddev).any() | torch.isnan(mu).any()):
            # torch.distributions.Normal.cdf validates its argument (acquisitions.py:21)
            raise ValueError("Expected value argument to be within the support of Normal: found NaN")
        dt = torch.float32 if mu.dtype == torch.float32 else torch.float64
        m = mu.detach().to(device=dev, dtype=dt).contiguous()
        s = stddev.detach().to(device=dev, dtype=dt).expand_as(m).contiguous()
        out = torch.empty_like(m)
        a = self.descriptor()
        with torch.cuda.device(dev):
            _lib.check(lib.bbo_acq_eval(C.byref(a), _lib.ptr(m), _lib.ptr(s), _lib.ptr(out), m.numel(),
                                        int(dt == torch.float32), _lib.stream_ptr(dev)), "bbo_acq_eval")
        return out.to(mu.dtype)


class EI(_DeviceAcq):
    """acquisitions.py:13-23: imp*Phi(Z) + sigma*phi(Z), imp = mu - best_f - eps, Z = imp/sigma."""
    kind = _lib.ACQ_EI

    def __init__(self, best_f, eps: float = 1e-6):
        adopt_reference_abcs(type(self), "AcquisitionFunction")
        self._best_f = float(best_f)
        self._eps = float(eps)

    def descriptor(self):
        return _lib.Acq(kind=self.kind, best_f=self._best_f, eps=self._eps, beta=0.0)


class UCB(_DeviceAcq):
    """acquisitions.py:26-31: mu + beta * sigma, beta = 1.8."""
    kind = _lib.ACQ_UCB

    def __init__(self, beta: float = 1.8):
        adopt_reference_abcs(type(self), "AcquisitionFunction")
        self.beta = float(beta)

    def descriptor(self):
        return _lib.Acq(kind=self.kind, best_f=0.0, eps=0.0, beta=self.beta)


class PI(_DeviceAcq):
    """Probability of improvement Phi((mu - best_f - eps)/sigma); not in the reference."""
    kind = _lib.ACQ_PI

    def __init__(self, best_f, eps: float = 1e-6):
        adopt_reference_abcs(type(self), "AcquisitionFunction")
        self._best_f = float(best_f)
        self._eps = float(eps)

    def descriptor(self):
        return _lib.Acq(kind=self.kind, best_f=self._best_f, eps=self._eps, beta=0.0)


def as_descriptor(acq) -> _lib.Acq:
    """Descriptor for this package's classes, or for the reference's own EI / UCB instances
    (duck-typed on their attrs fields: EI._best_f/_eps, UCB.beta)."""
    if isinstance(acq, _lib.Acq):
        return acq
    if isinstance(acq, _DeviceAcq):
        return acq.descriptor()
    name = type(acq).__name__
    if name == "EI" and hasattr(acq, "_best_f"):
        return _lib.Acq(kind=_lib.ACQ_EI, best_f=float(acq._best_f), eps=float(acq._eps), beta=0.0)
    if name == "UCB" and hasattr(acq, "beta"):
        return _lib.Acq(kind=_lib.ACQ_UCB, best_f=0.0, eps=0.0, beta=float(acq.beta))
    raise TypeError(f"unsupported acquisition function {name}")
