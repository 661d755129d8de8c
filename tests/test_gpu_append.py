"""GPU: incremental update of the factorised state (bbo_gp_append).

The appended state must equal the state a from-scratch factorisation of the enlarged data set gives with
the same hyper-parameters, and its posterior must match the CPU oracle on all n+m observations."""
import numpy as np
import pytest
import torch

from _golden import O
from _gpu_helpers import gp_from_params, np_, synth_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n0,m,d,kind", [
    (290, 10, 5, O.MATERN52),     # stays inside the padded size (384)
    (250, 12, 4, O.RBF),          # crosses a 128-row padding boundary (256 -> 384)
    (127, 1, 3, O.MATERN52),      # one point, fills the padding exactly
    (128, 1, 3, O.RBF),           # one point, first row of a new padding block
    (600, 70, 8, O.MATERN52),     # more than BBO_APPEND_MAX points: several library calls
])
def test_append_equals_refactorisation_and_oracle(n0, m, d, kind, cuda_device):
    p, X, y = synth_problem(100 + n0, n0 + m, d, kind=kind)
    gp = gp_from_params(p, X[:n0], y[:n0])
    gp.state                                          # factorise the first n0 observations
    gp.append(torch.as_tensor(X[n0:]), torch.as_tensor(y[n0:]).reshape(-1, 1))
    assert gp.n == n0 + m
    full = gp_from_params(p, X, y)
    sa, sf = gp.state, full.state
    n = n0 + m
    La, Lf = np_(sa["linv"])[:n, :n], np_(sf["linv"])[:n, :n]
    assert np.allclose(La, Lf, rtol=1e-9, atol=1e-9 * np.abs(Lf).max())
    assert np.allclose(np_(sa["alpha"])[:n], np_(sf["alpha"])[:n], rtol=1e-8, atol=1e-8 * np.abs(np_(sf["alpha"])).max())
    assert np.allclose(np_(sa["logdet"]), np_(sf["logdet"]), rtol=1e-10)
    # padding keeps the layout contract (unit diagonal, zeros elsewhere)
    pad = np_(sa["linv"])[n:, :]
    assert np.array_equal(pad[:, n:], np.eye(pad.shape[0])) and not pad[:, :n].any()
    assert not np_(sa["alpha"])[n:].any()
    # posterior and selection against the oracle on all observations
    rng = np.random.default_rng(7)
    Xq = rng.random((2000, d))
    best_f = float(y.max())
    import bbo_b200 as B
    top_s, top_i, mu, var, sc = gp.score_pool(torch.as_tensor(Xq).cuda(), B.EI(best_f), 5, return_all=True)
    oi, os_, omu, ovar, osc = O.score_pool(p, X, y, Xq, "ei", best_f, k=5)
    assert np.allclose(np_(mu), omu.numpy(), rtol=1e-6, atol=1e-8)
    assert np.allclose(np_(var), ovar.numpy(), rtol=1e-6, atol=1e-9)
    assert np.array_equal(np_(top_i), oi)
    # the TF32 mode picks up the new state as well
    ts, ti = gp.score_pool(torch.as_tensor(Xq).cuda(), B.EI(best_f), 1, precision="tf32")
    # (a state check, not an accuracy gate: the d = 3..5 problems here have cond(K) ~ 1e6, where the 3xTF32 K* error
    # times sum |alpha| ~ 1e5 is itself a few 1e-3 of the score; tests/test_gpu_tf32.py holds the 1e-3 gate)
    assert abs(float(ts[0]) - float(top_s[0])) <= 1e-2 * max(abs(float(top_s[0])), 1e-3)


def test_append_with_relabelled_targets_and_validation(cuda_device):
    p, X, y = synth_problem(5, 80, 3)
    gp = gp_from_params(p, X[:70], y[:70])
    y2 = (y - 0.3) * 1.7                                # what re-standardisation does to the old labels
    gp.append(torch.as_tensor(X[70:]), torch.as_tensor(y2[70:]).reshape(-1, 1),
              Y_all=torch.as_tensor(y2).reshape(-1, 1))
    full = gp_from_params(p, X, y2)
    assert np.allclose(np_(gp.state["alpha"]), np_(full.state["alpha"]), rtol=1e-8, atol=1e-9)
    with pytest.raises(ValueError):
        gp.append(torch.zeros(2, 4), torch.zeros(2, 1))
    with pytest.raises(ValueError):
        gp.append(torch.zeros(2, 3), torch.zeros(3, 1))
    gp.append(torch.zeros(0, 3), torch.zeros(0, 1))     # no-op
    assert gp.n == 80


def test_append_non_finite_point_raises_like_cholesky(cuda_device):
    p, X, y = synth_problem(9, 40, 2)
    gp = gp_from_params(p, X, y)
    gp.state
    bad = torch.full((1, 2), float("nan"), dtype=torch.float64)
    with pytest.raises(torch.linalg.LinAlgError):     # NaN pivot: what torch.linalg.cholesky reports (objectives.py:29)
        gp.append(bad, torch.zeros(1, 1, dtype=torch.float64))


def test_designer_incremental_and_warm_start(cuda_device):
    from bbo_b200.benchmarks import Branin2D
    from bbo_b200.designer import B200BODesigner
    f = Branin2D()
    des = B200BODesigner(f.bounds, "minimize", num_init=10, seed=3, warm_start=True, refit_every=4)
    fits = 0
    last = None
    for it in range(16):
        X = des.suggest(2)
        des.update(X, f(X))
        if des.last_surrogate is not None and des.last_surrogate is not last:
            fits += 1
            last = des.last_surrogate
    assert des._y.min() < 2.0, des._y.min()
    assert 2 <= fits <= 4, fits          # 11 model-guided suggestions, one fit per 4
    assert des.last_surrogate.n == len(des._y) - 2
