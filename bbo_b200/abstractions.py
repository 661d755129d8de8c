"""Plug-in interfaces of the hot path.

When the reference package ``bbo`` is importable its own ABCs are re-exported, so the
classes in this package *are* ``bbo`` surrogates / acquisition functions / optimizers.
Otherwise (e.g. on the GPU box, where the reference is absent) structurally identical
mirrors are defined:

  Surrogate               bbo/algorithms/abstractions.py:78-85
  AcquisitionFunction     bbo/algorithms/designers/bo/acquisitions.py:7-10
  GradientFreeOptimizer   bbo/algorithms/optimizers/base.py:8-18
"""
import abc

try:  # pragma: no cover - depends on the environment
    from bbo.algorithms.abstractions import Surrogate  # type: ignore
    from bbo.algorithms.optimizers.base import GradientFreeOptimizer  # type: ignore
    HAVE_BBO = True
except Exception:  # noqa: BLE001 - the reference may be absent or its optional deps missing
    HAVE_BBO = False

    class Surrogate(abc.ABC):
        @abc.abstractmethod
        def train(self, X=None, Y=None):
            pass

        @abc.abstractmethod
        def predict(self, X):
            pass

    class GradientFreeOptimizer(abc.ABC):
        @abc.abstractmethod
        def optimize(self, score_fn, problem_statement, *, count: int = 1, seed_candidates=tuple()):
            pass

try:  # the designers package imports gpytorch, which is usually missing
    from bbo.algorithms.designers.bo.acquisitions import AcquisitionFunction  # type: ignore
except Exception:  # noqa: BLE001
    class AcquisitionFunction(abc.ABC):
        @abc.abstractmethod
        def __call__(self, mu, stddev):
            pass
