"""Synthetic objectives used as workloads (numpy; cheap, not on the hot path).

Branin2D and Ackley follow bbo/benchmarks/experimenters/synthetic.py:47-93 (same constants,
bounds, optima); Rastrigin and Levy are named by BASELINE.json configs but absent from the
reference and follow the standard sfu.ca/~ssurjano definitions.  All are
MINIMIZE problems (``goal``); `unit_to_domain` maps the GP's [0,1]^d features to the native box.
When the reference is importable every class here IS a ``bbo`` ``SyntheticFunction``, so
``SyntheticExperimenter(Rastrigin(50))`` / ``SyntheticExperimenter(Levy(100))`` work unchanged.
"""
from __future__ import annotations

import numpy as np


try:  # the reference's own base class and goal enum when `bbo` is importable (benchmarks/experimenters/synthetic.py:13-24)
    from bbo.benchmarks.experimenters.synthetic import SyntheticFunction as _RefSyntheticFunction  # type: ignore
    from bbo.shared.base_study_config import ObjectiveMetricGoal  # type: ignore
    HAVE_BBO = True
except Exception:  # noqa: BLE001
    import abc
    import enum
    HAVE_BBO = False

    class ObjectiveMetricGoal(enum.Enum):     # shared/base_study_config.py:24-34
        MAXIMIZE = 1
        MINIMIZE = 2

        @property
        def is_maximize(self) -> bool:
            return self == self.MAXIMIZE

        @property
        def is_minimize(self) -> bool:
            return self == self.MINIMIZE

    class _RefSyntheticFunction(abc.ABC):     # synthetic.py:13-24
        num_objectives: int = 1

        @abc.abstractmethod
        def __call__(self, X):
            pass


class SyntheticFunction(_RefSyntheticFunction):
    """A reference ``SyntheticFunction`` (usable with ``SyntheticExperimenter(func)``, synthetic.py:27-45)
    plus the unit-cube helpers the tensor path uses."""
    dim: int
    bounds: list
    goal = ObjectiveMetricGoal.MINIMIZE
    optimal_value: float
    optimal_points = None

    def __new__(cls, *args, **kwargs):
        # import-order independence: if `bbo` was imported after this module, the mirror above is the base class;
        # register with the reference's ABC so SyntheticExperimenter's instance_of validator accepts the instance
        import sys
        mod = sys.modules.get("bbo.benchmarks.experimenters.synthetic")
        ref = getattr(mod, "SyntheticFunction", None) if mod is not None else None
        if ref is not None and not issubclass(cls, ref):
            ref.register(cls)
            goal = getattr(sys.modules.get("bbo.shared.base_study_config"), "ObjectiveMetricGoal", None)
            if goal is not None and not isinstance(cls.goal, goal):
                cls.goal = goal[cls.goal.name]
        return super().__new__(cls)

    def unit_to_domain(self, U: np.ndarray) -> np.ndarray:
        lb = np.array([b[0] for b in self.bounds], dtype=np.float64)
        ub = np.array([b[1] for b in self.bounds], dtype=np.float64)
        return lb + np.asarray(U, dtype=np.float64) * (ub - lb)

    def on_unit_cube(self, U: np.ndarray) -> np.ndarray:
        return self(self.unit_to_domain(U))


class Branin2D(SyntheticFunction):
    dim = 2
    bounds = [(-5, 10), (0, 15)]
    optimal_value = 0.397887
    optimal_points = [(-np.pi, 12.275), (np.pi, 2.275), (9.42478, 2.475)]

    def __call__(self, X):
        x1, x2 = X[..., 0], X[..., 1]
        b, c, t = 5.1 / (4 * np.pi ** 2), 5 / np.pi, 1 / (8 * np.pi)
        return (x2 - b * x1 ** 2 + c * x1 - 6) ** 2 + 10 * (1 - t) * np.cos(x1) + 10


class Ackley(SyntheticFunction):
    optimal_value = 0.0

    def __init__(self, dim, bounds=None):
        self.dim = dim
        self.bounds = bounds if bounds is not None else [(-32.768, 32.768)] * dim
        self.optimal_points = [tuple(0.0 for _ in range(dim))]

    def __call__(self, X):
        a, b, c = 20, 0.2, 2 * np.pi
        p1 = -a * np.exp(-np.linalg.norm(X, axis=-1) * b / np.sqrt(self.dim))
        p2 = -np.exp(np.mean(np.cos(c * X), axis=-1))
        return p1 + p2 + a + np.exp(1)


class Rastrigin(SyntheticFunction):
    optimal_value = 0.0

    def __init__(self, dim, bounds=None):
        self.dim = dim
        self.bounds = bounds if bounds is not None else [(-5.12, 5.12)] * dim
        self.optimal_points = [tuple(0.0 for _ in range(dim))]

    def __call__(self, X):
        return 10.0 * self.dim + np.sum(X ** 2 - 10.0 * np.cos(2 * np.pi * X), axis=-1)


class Levy(SyntheticFunction):
    optimal_value = 0.0

    def __init__(self, dim, bounds=None):
        self.dim = dim
        self.bounds = bounds if bounds is not None else [(-10.0, 10.0)] * dim
        self.optimal_points = [tuple(1.0 for _ in range(dim))]

    def __call__(self, X):
        w = 1.0 + (X - 1.0) / 4.0
        t1 = np.sin(np.pi * w[..., 0]) ** 2
        wi = w[..., :-1]
        t2 = np.sum((wi - 1) ** 2 * (1 + 10 * np.sin(np.pi * wi + 1) ** 2), axis=-1)
        wd = w[..., -1]
        t3 = (wd - 1) ** 2 * (1 + np.sin(2 * np.pi * wd) ** 2)
        return t1 + t2 + t3


def make_observations(func: SyntheticFunction, n: int, seed: int):
    """Seeded training set in the designer's conventions: features in [0,1]^d,
    labels sign-flipped for MINIMIZE (bo.py:80-83) and standardised (bo.py:126-128)."""
    rng = np.random.default_rng(seed)
    X = rng.random((n, func.dim))
    y = -func.on_unit_cube(X)
    y = (y - y.mean()) / (y.std(ddof=1) + 1e-6)
    return X, y
