"""B200BODesigner: the reference's BODesigner ask/tell loop (bbo/algorithms/designers/bo/bo.py:36-128)
on arrays, with the whole suggest() hot path on the GPU.

  suggest(count)  bo.py:95-118: random init for the first num_init points; then
                  to_xy -> [0,1] features, sign-flipped standardised labels (bo.py:123-128,
                  converters/core.py:406-407) -> surrogate_factory(X, Y, device) -> train() ->
                  acquisition_function_factory(Y) (default EI(Y.max()), bo.py:54-57) ->
                  acquisition optimiser -> best candidates.
  update(X, y)    bo.py:120-121.

Differences from the reference, all deliberate: operates on numpy arrays instead of ``Trial``s (the
data model is out of scope; ``INTEGRATION.md`` shows how to plug the same pieces into the real
``BODesigner``); the default surrogate is ``CudaGP`` (the reference's default ``GPyTorchGP`` is
third-party) trained with ``sign=-1`` (maximising the likelihood, as GPyTorchGP does);
``suggest(count)`` honours ``count`` (the reference ignores it after initialisation, bo.py:117).

Opt-in extensions (the defaults keep the reference's refit-from-scratch behaviour):
  ``warm_start=True``   the new surrogate starts Adam from the previous suggestion's hyper-parameters
                        instead of the factory's initial values (bo.py:105 builds a fresh model each time).
  ``refit_every=k``     hyper-parameters are refitted only at every k-th suggestion; in between, new
                        observations are appended to the factorised state (``CudaGP.append``: O(m n^2)
                        instead of epochs x O(n^3)).
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np
import torch

from .acquisitions import EI
from .gp import CudaGP
from .kernels import ConstantMean, RBFKernel, ScaleKernel
from .optimizers import PoolOptimizer, TensorScoreFn


def default_surrogate_factory(train_X: torch.Tensor, train_Y: torch.Tensor, device):
    cov = ScaleKernel(RBFKernel(ard_dim=train_X.shape[1])).double()
    with torch.no_grad():   # deterministic start instead of the reference's N(0,1) draw (kernels.py:31-34,51)
        cov.base_kernel.lengthscale.fill_(0.0)
        cov.variance.fill_(0.5413)
    return CudaGP(train_X, train_Y, ConstantMean().double(), cov, device=device, noise_variance=1e-4,
                  lr=0.05, epochs=60, sign=-1.0)


def _copy_parameters(src, dst) -> None:
    """Warm start: previous fit's raw mean / kernel parameters -> the new surrogate's modules."""
    for a, b in ((src._mean_module, dst._mean_module), (src._cov_module, dst._cov_module)):
        pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
        with torch.no_grad():
            for name, p in pb.items():
                if name in pa and pa[name].shape == p.shape:
                    p.copy_(pa[name].to(p.dtype))


class B200BODesigner:
    def __init__(self, bounds, goal: str = "minimize", *, num_init: int = 10, seed: Optional[int] = None,
                 device="cuda", surrogate_factory: Callable = default_surrogate_factory,
                 acquisition_function_factory: Callable = lambda Y: EI(float(Y.max())),
                 acquisition_optimizer_factory: Callable = lambda rng: PoolOptimizer(8192, int(rng.integers(1 << 31))),
                 precision: str = "fp64", warm_start: bool = False, refit_every: int = 1):
        self.bounds = np.asarray(bounds, dtype=np.float64).reshape(-1, 2)
        self.d = self.bounds.shape[0]
        if goal not in ("minimize", "maximize"):
            raise ValueError("goal must be 'minimize' or 'maximize'")
        self.goal = goal
        self._num_init = int(num_init)
        self._device = device
        self._rng = np.random.default_rng(seed)
        self._surrogate_factory = surrogate_factory
        self._acqf_factory = acquisition_function_factory
        self._optimizer = acquisition_optimizer_factory(self._rng)
        self._precision = precision
        self._X = np.zeros((0, self.d))
        self._y = np.zeros((0,))
        self.last_surrogate = None
        if refit_every < 1:
            raise ValueError("refit_every must be >= 1")
        self._warm_start = bool(warm_start)
        self._refit_every = int(refit_every)
        self._since_fit = 0          # suggestions served by the current hyper-parameters
        self._n_in_surrogate = 0     # observations the current surrogate has seen

    # converters/core.py:211-216: features scaled to [0,1]
    def _to_unit(self, X):
        return (np.asarray(X, dtype=np.float64) - self.bounds[:, 0]) / (self.bounds[:, 1] - self.bounds[:, 0])

    def _from_unit(self, U):
        return self.bounds[:, 0] + np.asarray(U, dtype=np.float64) * (self.bounds[:, 1] - self.bounds[:, 0])

    def suggest(self, count: Optional[int] = None) -> np.ndarray:
        count = count or 1
        if len(self._y) < self._num_init:
            return self._from_unit(self._rng.random((count, self.d)))
        X = torch.as_tensor(self._to_unit(self._X))
        y = torch.as_tensor(self._y if self.goal == "maximize" else -self._y).reshape(-1, 1)   # bo.py:80-83
        Y = (y - y.mean(dim=0)) / (y.std(dim=0) + 1e-6)                                      # bo.py:126-128
        prev = self.last_surrogate
        n_seen = self._n_in_surrogate
        if (prev is not None and hasattr(prev, "append") and self._since_fit + 1 < self._refit_every
                and len(self._y) >= n_seen):
            surrogate = prev                     # same hyper-parameters, enlarged data set
            surrogate.append(X[n_seen:], Y[n_seen:], Y_all=Y)
            self._since_fit += 1
        else:
            surrogate = self._surrogate_factory(X, Y, self._device)
            if self._warm_start and prev is not None:
                _copy_parameters(prev, surrogate)
            surrogate.train()
            self._since_fit = 0
        self._n_in_surrogate = len(self._y)
        acqf = self._acqf_factory(Y)
        score_fn = TensorScoreFn(surrogate, acqf, self._precision)
        best, scores, _ = self._optimizer.optimize_tensor(score_fn, self.d, count=count)
        self.last_surrogate = surrogate
        return self._from_unit(best.cpu().numpy())

    def update(self, X, y) -> None:
        X = np.asarray(X, dtype=np.float64).reshape(-1, self.d)
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        if len(X) != len(y):
            raise ValueError("X and y must have the same number of rows")
        self._X = np.concatenate([self._X, X])
        self._y = np.concatenate([self._y, y])
