"""PoolOptimizer: the acquisition optimiser of the hot path.

Implements the reference's ``GradientFreeOptimizer.optimize(score_fn, problem_statement, *, count,
seed_candidates)`` contract (bbo/algorithms/optimizers/base.py:8-18) and replaces
``DesignerAsOptimizer`` + ``RandomDesigner`` + ``get_best_trials``
(optimizers/designer_optimizer.py:18-43, designers/random.py:41-57, shared/trial.py:278-294):
a uniform candidate pool is generated on the device, scored by the fused GP/acquisition kernels and
reduced to the ``count`` best rows -- no Python ``Trial`` is created for a pool member.

Two ways in:
  * tensor fast path -- ``score_fn`` is a ``TensorScoreFn`` (surrogate + acquisition): everything
    stays on the GPU; used by ``B200BODesigner``.
  * reference path   -- ``score_fn`` is the closure built by the reference's ``BODesigner.suggest``
    (bo.py:110-115, Sequence[Trial] -> {'acquisition': tensor}); needs the ``bbo`` package for the
    Trial conversion.  The pool is scored in ONE call (batch = pool) instead of 100 calls of 1.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .abstractions import GradientFreeOptimizer
from .pool import uniform_pool


class TensorScoreFn:
    """What BODesigner.score_fn closes over (bo.py:110-115), kept as data so it can be fused."""

    def __init__(self, surrogate, acquisition, precision: str = "fp64"):
        self.surrogate = surrogate
        self.acquisition = acquisition
        self.precision = precision

    def __call__(self, X: torch.Tensor) -> torch.Tensor:      # (q,d) -> (q,) scores
        mu, var = self.surrogate.predict_diag(X)
        return self.acquisition(mu, var.sqrt())


class PoolOptimizer(GradientFreeOptimizer):
    def __init__(self, num_candidates: int = 8192, seed: int = 0, dtype=torch.float64):
        self.num_candidates = int(num_candidates)
        self.seed = int(seed)
        self.dtype = dtype
        self._calls = 0

    # -- tensor fast path -----------------------------------------------------------------
    def optimize_tensor(self, score_fn: TensorScoreFn, d: int, *, count: int = 1,
                        seed_candidates: Optional[torch.Tensor] = None):
        """Returns (X_best (count,d) float64 in [0,1]^d, scores (count,), global indices (count,))."""
        gp = score_fn.surrogate
        dev = gp._device
        self._calls += 1
        pool = uniform_pool(self.seed + 7919 * self._calls, 0, self.num_candidates, d, device=dev,
                            dtype=self.dtype)
        if seed_candidates is not None and len(seed_candidates):
            pool = torch.cat([seed_candidates.to(device=dev, dtype=self.dtype), pool])
        s, i = gp.score_pool(pool, score_fn.acquisition, count, precision=score_fn.precision)
        valid = i >= 0
        return pool[i[valid]].to(torch.float64), s[valid], i[valid]

    # -- reference contract ---------------------------------------------------------------
    def optimize(self, score_fn, problem_statement, *, count: int = 1, seed_candidates: Sequence = tuple()) -> List:
        if isinstance(score_fn, TensorScoreFn):
            d = problem_statement if isinstance(problem_statement, int) else len(problem_statement)
            return self.optimize_tensor(score_fn, d, count=count)[0]
        try:
            from bbo.algorithms.converters.torch_converter import TrialToTorchConverter  # type: ignore
            from bbo.shared.trial import Measurement, get_best_trials  # type: ignore
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("PoolOptimizer.optimize with a Trial-level score_fn needs the reference "
                               "package `bbo` on sys.path; use optimize_tensor otherwise") from e
        conv = TrialToTorchConverter.from_study_config(problem_statement)
        d = problem_statement.search_space.num_parameters()
        self._calls += 1
        dev = "cuda"
        pool = uniform_pool(self.seed + 7919 * self._calls, 0, self.num_candidates, d, device=dev,
                            dtype=torch.float32).cpu()
        trials = list(seed_candidates) + _features_to_trials(conv, pool)
        with torch.no_grad():
            scores = score_fn(trials)                 # ONE batched call: (q,1,d) -> CudaGP fast path
        name = problem_statement.metric_information_item().name
        vals = scores[name].reshape(-1).tolist()
        for t, v in zip(trials, vals):
            t.complete(Measurement({name: v}))
        return get_best_trials(trials, problem_statement, count=count)


def _features_to_trials(conv, pool: torch.Tensor):
    """[0,1]^d rows -> Trials through the reference converter (torch_converter.py:57-60)."""
    from bbo.algorithms.converters.torch_converter import TypeArray  # type: ignore
    feats = TypeArray(pool.to(torch.float32), torch.zeros(pool.shape[0], 0, dtype=torch.int32))
    return conv.to_trials(feats)
