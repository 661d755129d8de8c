"""Host logic of the drop-in, on the CPU: the UNMODIFIED reference BODesigner (designers/bo/bo.py:36-128) runs
ask/tell rounds with PoolOptimizer as its acquisition optimiser (Trial-level path: TypeArray construction,
one batched score_fn call, get_best_trials).  The device pool generator is replaced by torch.rand here (no GPU
in this suite); the surrogate is the reference's own GP behind the (q,1,d) adapter."""
import numpy as np
import pytest
import torch

from _refenv import RefGPAdapter, register_adapter, require_reference


@pytest.fixture()
def ref():
    require_reference()
    register_adapter()
    import bbo_b200.optimizers as opt

    def cpu_pool(seed, row_offset, q, d, *, device="cpu", dtype=torch.float32, out=None):
        g = torch.Generator().manual_seed(int(seed) & 0x7fffffff)
        return torch.rand(q, d, generator=g, dtype=dtype)

    old = opt.uniform_pool
    opt.uniform_pool = cpu_pool
    yield opt
    opt.uniform_pool = old


def _branin_experimenter():
    from bbo.benchmarks.experimenters.synthetic import SyntheticExperimenter
    from bbo_b200.benchmarks import Branin2D
    return SyntheticExperimenter(Branin2D())


def test_pool_optimizer_is_a_reference_optimizer(ref):
    from bbo.algorithms.optimizers.base import GradientFreeOptimizer
    assert isinstance(ref.PoolOptimizer(16), GradientFreeOptimizer)


def test_features_to_trials_roundtrip(ref):
    from bbo.algorithms.converters.torch_converter import TrialToTorchConverter
    ps = _branin_experimenter().problem_statement()
    conv = TrialToTorchConverter.from_study_config(ps)
    pool = torch.rand(33, 2, dtype=torch.float32)
    trials = ref._features_to_trials(conv, pool)
    assert len(trials) == 33
    back = conv.to_features(trials).double
    assert back.shape == (33, 2)
    assert torch.allclose(back, pool, atol=1e-6)
    x0 = trials[0].parameters["x0"].value
    assert -5.0 <= x0 <= 10.0


def test_closure_parts_reads_the_reference_closure(ref):
    surrogate, acqf_fn = object(), object()

    def score_fn(trials):
        return surrogate, acqf_fn
    s, a = ref.closure_parts(score_fn)
    assert s is surrogate and a is acqf_fn
    assert ref.closure_parts(lambda t: t) == (None, None)


def test_unmodified_bodesigner_ask_tell_with_pool_optimizer(ref):
    from bbo.algorithms.abstractions import CompletedTrials
    from bbo.algorithms.designers.bo.acquisitions import EI
    from bbo.algorithms.designers.bo.bo import BODesigner
    from bbo.algorithms.surrogates.gp import GP
    from bbo.algorithms.surrogates.gp.kernels import RBFKernel, ScaleKernel
    from bbo.algorithms.surrogates.gp.means import ConstantMean

    exp = _branin_experimenter()
    ps = exp.problem_statement()
    torch.manual_seed(0)

    def surrogate_factory(X, Y, device):
        gp = GP(X.double(), Y.double(), ConstantMean().double(), ScaleKernel(RBFKernel()).double(),
                epochs=5, noise_variance=1e-4)
        return RefGPAdapter(gp)

    opts = []

    def optimizer_factory(rng):
        opts.append(ref.PoolOptimizer(256, seed=3))
        return opts[-1]

    des = BODesigner(ps, num_init=6, np_seed_or_rng=0, surrogate_factory=surrogate_factory,
                     acquisition_function_factory=lambda Y: EI(Y.max()),
                     acquisition_optimizer_factory=optimizer_factory)
    seen = []
    for it in range(9):
        trials = des.suggest()
        assert len(trials) >= 1
        exp.evaluate(trials)
        des.update(CompletedTrials(trials))
        seen.extend(t.final_measurement.metrics["y"].value for t in trials)
    assert opts[0].last_path == "closure-trials"      # the adapter has no score_pool: Trial-level path
    assert len(seen) == 9 and np.all(np.isfinite(seen))


def test_trial_path_selection_is_get_best_trials(ref):
    """count > 1, ties and seed candidates: same answer as the reference's own selection."""
    from bbo.shared.trial import get_best_trials
    ps = _branin_experimenter().problem_statement()
    import copy
    from bbo.shared.base_study_config import MetricInformation, ObjectiveMetricGoal
    acq_ps = copy.deepcopy(ps)
    acq_ps.metric_information = [MetricInformation("acquisition", ObjectiveMetricGoal.MAXIMIZE)]
    po = ref.PoolOptimizer(64, seed=1, fuse_closure=False)
    calls = []

    def score_fn(trials):
        calls.append(len(trials))
        v = torch.tensor([round(t.parameters["x0"].value) % 5 for t in trials], dtype=torch.float32)
        return {"acquisition": v.reshape(-1, 1, 1)}

    best = po.optimize(score_fn, acq_ps, count=4)
    assert calls == [64]                               # ONE batched call
    assert len(best) == 4
    vals = [t.final_measurement.metrics["acquisition"].value for t in best]
    assert vals == sorted(vals, reverse=True)
