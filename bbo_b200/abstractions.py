"""Plug-in interfaces of the hot path.

When the reference package ``bbo`` is importable (already on ``sys.path``, or under ``$BBO_REFERENCE``) its
own ABCs are the base classes, so the classes in this package *are* ``bbo`` surrogates / acquisition
functions / optimizers.  Otherwise structurally identical mirrors are defined:

  Surrogate               bbo/algorithms/abstractions.py:78-85
  AcquisitionFunction     bbo/algorithms/designers/bo/acquisitions.py:7-10
  GradientFreeOptimizer   bbo/algorithms/optimizers/base.py:8-18

If ``bbo`` only becomes importable AFTER this package was imported (the mirrors are then the bases),
``adopt_reference_abcs`` registers the classes as virtual subclasses of the reference ABCs, so that
``BODesigner``'s ``instance_of(GradientFreeOptimizer)`` validator (bo.py:64) accepts them either way.
It is called from the constructors of ``PoolOptimizer`` / ``CudaGP`` / the acquisition classes.
"""
import abc
import os
import sys


def _reference_on_path() -> None:
    p = os.environ.get("BBO_REFERENCE")
    if p and os.path.isdir(os.path.join(p, "bbo")) and p not in sys.path:
        sys.path.insert(0, p)


_reference_on_path()

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
    # a failed package import leaves its already-imported children in sys.modules; a later, successful import of
    # bbo.algorithms.designers (e.g. once gpytorch or its stub is available) would then trip over them
    for _k in [k for k in sys.modules if k.startswith("bbo.algorithms.designers")]:
        sys.modules.pop(_k, None)

    class AcquisitionFunction(abc.ABC):
        @abc.abstractmethod
        def __call__(self, mu, stddev):
            pass


_REFERENCE_ABCS = {
    "Surrogate": ("bbo.algorithms.abstractions", "Surrogate"),
    "GradientFreeOptimizer": ("bbo.algorithms.optimizers.base", "GradientFreeOptimizer"),
    "AcquisitionFunction": ("bbo.algorithms.designers.bo.acquisitions", "AcquisitionFunction"),
}


def adopt_reference_abcs(cls, role: str) -> None:
    """Register ``cls`` with the reference's ABC for ``role`` if that module is loaded and ``cls`` does not
    already derive from it (import-order independence; no import is triggered here)."""
    modname, attr = _REFERENCE_ABCS[role]
    mod = sys.modules.get(modname)
    ref = getattr(mod, attr, None) if mod is not None else None
    if ref is not None and isinstance(ref, abc.ABCMeta) and not issubclass(cls, ref):
        ref.register(cls)
