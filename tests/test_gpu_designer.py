"""GPU: the ask/tell loop on Branin-2D (BASELINE config 1 shape) and the PoolOptimizer tensor path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_branin_bo_loop_improves(cuda_device):
    from bbo_b200.benchmarks import Branin2D
    from bbo_b200.designer import B200BODesigner
    f = Branin2D()
    des = B200BODesigner(f.bounds, "minimize", num_init=10, seed=0)
    rng_best = []
    for it in range(14):
        X = des.suggest(2 if it >= 5 else 2)
        assert X.shape == (2, 2)
        assert (X >= np.array([-5, 0]) - 1e-9).all() and (X <= np.array([10, 15]) + 1e-9).all()
        des.update(X, f(X))
        rng_best.append(des._y.min())
    # after 10 random + 18 model-guided evaluations BO must be close to the optimum 0.3979
    assert des._y.min() < 2.0, des._y.min()
    assert des._y.min() <= rng_best[4] + 1e-12
    assert des.last_surrogate is not None and des.last_surrogate.last_values is not None
    v = des.last_surrogate.last_values.cpu().numpy()
    assert v[-1] > v[0]          # sign=-1: the likelihood went up during the fit


def test_pool_optimizer_matches_manual_scoring(cuda_device):
    import bbo_b200 as B
    from bbo_b200.optimizers import PoolOptimizer, TensorScoreFn
    from bbo_b200.pool import uniform_pool
    from _gpu_helpers import gp_from_params, synth_problem
    p, X, y = synth_problem(21, 100, 2)
    gp = gp_from_params(p, X, y)
    acq = B.EI(float(y.max()))
    opt = PoolOptimizer(8192, seed=5)
    best, scores, idx = opt.optimize_tensor(TensorScoreFn(gp, acq), 2, count=3)
    pool = uniform_pool(5 + 7919, 0, 8192, 2, device="cuda")
    sc = TensorScoreFn(gp, acq)(pool).cpu().numpy()
    order = np.lexsort((np.arange(8192), -sc))[:3]
    np.testing.assert_array_equal(idx.cpu().numpy(), order)
    np.testing.assert_allclose(best.cpu().numpy(), pool[order].cpu().numpy())
    np.testing.assert_allclose(scores.cpu().numpy(), sc[order], rtol=1e-12)


def test_objectives_known_optima():
    """bbo/benchmarks/experimenters/synthetic_test.py:8-40 style known-answer checks (CPU math)."""
    from bbo_b200.benchmarks import Ackley, Branin2D, Levy, Rastrigin
    b = Branin2D()
    for pt in b.optimal_points:
        assert abs(b(np.array(pt)) - b.optimal_value) < 1e-4
    for f in (Ackley(20), Rastrigin(50), Levy(100)):
        assert abs(f(np.array(f.optimal_points[0])) - f.optimal_value) < 1e-9
        assert f(f.unit_to_domain(np.full(f.dim, 0.9))) > 0.1


def test_evolution_multistart_beats_its_own_random_generation(cuda_device):
    """64 RE1 populations (pop 25, tournament 5, 1 mutation, 100 evaluations each, bo.py:26-33 defaults)
    scored through the fused kernels: the result is the best of everything evaluated and improves on
    the random first generation."""
    import bbo_b200 as B
    from bbo_b200.evolution import EvolutionMultiStart, select_topk
    from bbo_b200.optimizers import TensorScoreFn
    from _gpu_helpers import gp_from_params, synth_problem
    p, X, y = synth_problem(31, 300, 6)
    gp = gp_from_params(p, X, y)
    fn = TensorScoreFn(gp, B.UCB(2.0))
    calls = []
    def counted(Xq):
        calls.append(Xq.shape[0])
        return fn(Xq)
    counted.surrogate = gp
    opt = EvolutionMultiStart(num_restarts=64, num_evaluations=100, seed=3)
    best, scores, idx = opt.optimize_tensor(counted, 6, count=5)
    assert best.shape == (5, 6) and (best >= 0).all() and (best < 1).all()
    assert calls[0] == 64 * 25 and calls[1:] == [64] * 75          # one fused call per generation
    assert (scores[:-1] >= scores[1:]).all()
    np.testing.assert_allclose(fn(best).cpu().numpy(), scores.cpu().numpy(), rtol=1e-12)
    gen0_best = fn(torch.rand(64 * 25, 6, dtype=torch.float64, device="cuda")).max()
    assert float(scores[0]) >= float(gen0_best) - 0.05 * abs(float(gen0_best))
    # select_topk == stable descending order, also across >2048 entries
    s = torch.round(torch.randn(5000, dtype=torch.float64, device="cuda"), decimals=1)
    got = select_topk(s, 17).cpu().numpy()
    want = np.lexsort((np.arange(5000), -s.cpu().numpy()))[:17]
    np.testing.assert_array_equal(got, want)
