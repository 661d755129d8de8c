"""PoolOptimizer: the acquisition optimiser of the hot path.

Implements the reference's ``GradientFreeOptimizer.optimize(score_fn, problem_statement, *, count,
seed_candidates)`` contract (bbo/algorithms/optimizers/base.py:8-18) and replaces
``DesignerAsOptimizer`` + ``RandomDesigner`` + ``get_best_trials``
(optimizers/designer_optimizer.py:18-43, designers/random.py:41-57, shared/trial.py:278-294):
a uniform candidate pool is generated on the device, scored by the fused GP/acquisition kernels and
reduced to the ``count`` best rows.

Three ways in:
  * tensor fast path -- ``score_fn`` is a ``TensorScoreFn`` (surrogate + acquisition): everything
    stays on the GPU; used by ``B200BODesigner``.
  * reference closure, fused -- ``score_fn`` is the closure built by the UNMODIFIED reference
    ``BODesigner.suggest`` (bo.py:110-115).  Its free variables ``surrogate`` / ``acqf_fn`` are read from the
    closure cells; when the surrogate is a ``CudaGP`` and the acquisition is EI/UCB/PI (ours or the
    reference's), the pool is scored by ``CudaGP.score_pool`` on the device and only the ``count`` winners
    become ``Trial``s (``fuse_closure=True``, the default).
  * reference closure, Trial level -- any other ``score_fn`` (or ``fuse_closure=False``): the pool becomes
    Trials once through the reference converter and is scored in ONE batched ``score_fn`` call (instead of
    the reference's 100 calls of 1, designer_optimizer.py:29-41), then ``get_best_trials``.
Both closure paths need the reference package ``bbo`` for the Trial conversion.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .abstractions import GradientFreeOptimizer, adopt_reference_abcs
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


def closure_parts(score_fn):
    """(surrogate, acquisition) captured by the reference's score_fn closure (bo.py:110-115: free variables
    ``surrogate`` and ``acqf_fn``), or (None, None) if ``score_fn`` is not that closure."""
    code = getattr(score_fn, "__code__", None)
    cells = getattr(score_fn, "__closure__", None)
    if code is None or not cells:
        return None, None
    try:
        env = {name: cell.cell_contents for name, cell in zip(code.co_freevars, cells)}
    except ValueError:      # empty cell
        return None, None
    return env.get("surrogate"), env.get("acqf_fn")


class PoolOptimizer(GradientFreeOptimizer):
    def __init__(self, num_candidates: int = 8192, seed: int = 0, dtype=torch.float64, *,
                 precision: str = "fp64", fuse_closure: bool = True, device="cuda"):
        self.num_candidates = int(num_candidates)
        self.seed = int(seed)
        self.dtype = dtype
        self.precision = precision
        self.fuse_closure = bool(fuse_closure)
        self.device = device
        self._calls = 0
        self.last_path = None        # "tensor" | "closure-fused" | "closure-trials"
        adopt_reference_abcs(type(self), "GradientFreeOptimizer")

    def _pool(self, d: int, device, dtype) -> torch.Tensor:
        self._calls += 1
        return uniform_pool(self.seed + 7919 * self._calls, 0, self.num_candidates, d, device=device, dtype=dtype)

    # -- tensor fast path -----------------------------------------------------------------
    def optimize_tensor(self, score_fn: TensorScoreFn, d: int, *, count: int = 1,
                        seed_candidates: Optional[torch.Tensor] = None):
        """Returns (X_best (count,d) float64 in [0,1]^d, scores (count,), global indices (count,))."""
        gp = score_fn.surrogate
        dev = gp._device
        pool = self._pool(d, dev, self.dtype)
        if seed_candidates is not None and len(seed_candidates):
            pool = torch.cat([seed_candidates.to(device=dev, dtype=self.dtype), pool])
        s, i = gp.score_pool(pool, score_fn.acquisition, count, precision=score_fn.precision)
        valid = i >= 0
        self.last_path = "tensor"
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
        name = problem_statement.metric_information_item().name
        seeds = list(seed_candidates)

        surrogate, acqf = closure_parts(score_fn) if self.fuse_closure else (None, None)
        if surrogate is not None and hasattr(surrogate, "score_pool") and _describable(acqf):
            # float32 features in [0,1]^d, what the reference converter produces (torch_converter.py:66)
            pool = self._pool(d, surrogate._device, torch.float32)
            if seeds:
                sx = conv.to_features(seeds).double.to(device=pool.device, dtype=torch.float32)
                pool = torch.cat([sx, pool])
            k = min(int(count), pool.shape[0])
            with torch.no_grad():
                s, i = surrogate.score_pool(pool, acqf, k, precision=self.precision)
            valid = i >= 0
            rows, vals = pool[i[valid]].cpu(), s[valid].cpu().tolist()
            trials = _features_to_trials(conv, rows)
            for t, v in zip(trials, vals):
                t.complete(Measurement({name: v}))
            self.last_path = "closure-fused"
            return trials       # already in get_best_trials order: score descending, ties -> lowest index

        pool = self._pool(d, self.device, torch.float32).cpu()
        trials = seeds + _features_to_trials(conv, pool)
        with torch.no_grad():
            scores = score_fn(trials)                 # ONE batched call: (q,1,d) -> CudaGP.predict fast path
        vals = scores[name].reshape(-1).tolist()
        for t, v in zip(trials, vals):
            t.complete(Measurement({name: v}))
        self.last_path = "closure-trials"
        return get_best_trials(trials, problem_statement, count=count)


def _describable(acqf) -> bool:
    from .acquisitions import as_descriptor
    try:
        as_descriptor(acqf)
        return True
    except Exception:  # noqa: BLE001
        return False


def _features_to_trials(conv, pool: torch.Tensor):
    """[0,1]^d float32 rows -> Trials through the reference converter (torch_converter.py:45-60).  ``TypeArray``
    (converters/core.py:629-634) has four required fields; a BO search space is all-double (bo.py:72-74), so the
    other three are empty (q,0) arrays of the converter's dtypes."""
    from bbo.algorithms.converters.core import TypeArray  # type: ignore
    q = pool.shape[0]
    feats = TypeArray(double=pool.detach().to("cpu", torch.float32).contiguous(),
                      integer=torch.zeros(q, 0, dtype=torch.int32),
                      categorical=torch.zeros(q, 0, dtype=torch.int32),
                      discrete=torch.zeros(q, 0, dtype=torch.float32))
    return conv.to_trials(feats)
