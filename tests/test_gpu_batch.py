"""GPU parity of the batched small-problem path (BASELINE config 4 shape, scaled down for the oracle)."""
import numpy as np
import pytest
import torch

from _golden import O
from _gpu_helpers import np_, synth_problem

pytestmark = pytest.mark.gpu


def _batch(B, n, d, kind, seed):
    ps, Xs, ys = [], [], []
    for b in range(B):
        p, X, y = synth_problem(seed + b, n, d, kind=kind, noise=1e-4)
        ps.append(p), Xs.append(X), ys.append(y)
    theta = np.stack([np.concatenate([[p.constant], np.atleast_1d(p.raw_lengthscale), [p.raw_variance]]) for p in ps])
    return ps, np.stack(Xs), np.stack(ys), theta


@pytest.fixture()
def single_launch():
    """Force the single-launch small-problem kernel (k_fit_small; default only from 96 problems on)."""
    import os
    old = os.environ.get("BBO_FIT_SMALL_MIN_BATCH")
    os.environ["BBO_FIT_SMALL_MIN_BATCH"] = "1"
    yield
    if old is None:
        del os.environ["BBO_FIT_SMALL_MIN_BATCH"]
    else:
        os.environ["BBO_FIT_SMALL_MIN_BATCH"] = old


@pytest.mark.parametrize("B,n,d,kind", [(9, 256, 10, O.RBF), (12, 130, 7, O.MATERN52), (8, 500, 33, O.RBF),
                                        (2, 64, 3, O.MATERN52)])
def test_single_launch_value_grad_train_and_scoring(B, n, d, kind, cuda_device, single_launch):
    """The same oracle checks as below with every problem on one CTA of k_fit_small."""
    test_batched_value_grad_train_and_scoring(B, n, d, kind, cuda_device)


# the multi-kernel pipeline with a batch index (the default below 96 problems)
@pytest.mark.parametrize("B,n,d,kind", [(5, 70, 4, O.MATERN52), (3, 256, 10, O.RBF)])
def test_batched_value_grad_train_and_scoring(B, n, d, kind, cuda_device):
    import bbo_b200 as B200
    from bbo_b200.batch import CudaGPBatch
    ps, X, y, theta = _batch(B, n, d, kind, 300)
    impl = "rbf_vmap" if kind == O.RBF else "matern52_vmap"
    gb = CudaGPBatch(torch.as_tensor(X), torch.as_tensor(y)[..., None], impl, theta=torch.as_tensor(theta),
                     noise_variance=1e-4, epochs=6, lr=0.01)
    vals, grads = gb.value_and_grad()
    for b in range(B):
        oval, og_ls, og_var, og_c = O.objective_value_and_grad(ps[b], X[b], y[b])
        assert abs(float(vals[b]) - oval) <= 1e-7 * abs(oval)
        ls = O.softplus(ps[b].raw_lengthscale).numpy()
        # oracle gradients are w.r.t. raw parameters: undo the softplus chain rule
        g_l = og_ls / O.sigmoid(ps[b].raw_lengthscale).numpy()
        np.testing.assert_allclose(np_(grads[b, :d]), g_l, rtol=0, atol=1e-6 * np.abs(g_l).max())
        assert abs(float(grads[b, d]) - og_var / float(O.sigmoid(ps[b].raw_variance))) <= 1e-6 * abs(og_var) + 1e-9
        assert abs(float(grads[b, d + 1]) - og_c) <= 1e-6 * abs(og_c) + 1e-9
        assert ls.shape == (d,)
    # batched scoring before training == per-problem oracle
    q, k = 600, 4
    Xq = np.random.default_rng(1).random((B, q, d))
    s, i = gb.score_pools(torch.as_tensor(Xq), lambda b: B200.EI(float(y[b].max())), k)
    for b in range(B):
        oi, os_, *_ = O.score_pool(ps[b], X[b], y[b], Xq[b], "ei", float(y[b].max()), k=k)
        np.testing.assert_array_equal(np_(i[b]), oi)
        np.testing.assert_allclose(np_(s[b]), os_, rtol=1e-6, atol=1e-12)
    # batched Adam == per-problem oracle Adam (gp.py:54-63, sign=+1 as in the reference)
    gb.train()
    for b in range(B):
        final, ovals = O.adam_train(ps[b], X[b], y[b], lr=0.01, epochs=6)
        got = np_(gb.theta[b])
        want = np.concatenate([[final.constant], np.atleast_1d(final.raw_lengthscale), [final.raw_variance]])
        np.testing.assert_allclose(got, want, rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(np_(gb.last_values[:, b]), ovals, rtol=1e-7)
    assert int(gb.info.abs().sum()) == 0


def test_batched_scoring_many_problems_two_groups(cuda_device):
    """300 problems (two launch groups), pools spanning two top-k segments, per-problem incumbents: the batched
    library call against the per-problem path and the oracle."""
    import bbo_b200 as B200
    from bbo_b200.batch import CudaGPBatch
    B, n, d, q, k = 300, 64, 3, 2100, 16
    ps, X, y, theta = _batch(B, n, d, O.RBF, 900)
    gb = CudaGPBatch(torch.as_tensor(X), torch.as_tensor(y)[..., None], "rbf_vmap", theta=torch.as_tensor(theta),
                     noise_variance=1e-4)
    Xq = np.random.default_rng(2).random((B, q, d))
    best = [float(y[b].max()) - 0.1 * (b % 3) for b in range(B)]
    s, i = gb.score_pools(torch.as_tensor(Xq), lambda b: B200.EI(best[b]), k)
    assert s.shape == (B, k) and i.shape == (B, k)
    for b in (0, 1, 150, 239, 240, 299):                      # both sides of the group boundary
        oi, os_, *_ = O.score_pool(ps[b], X[b], y[b], Xq[b], "ei", best[b], k=k)
        np.testing.assert_array_equal(np_(i[b]), oi)
        np.testing.assert_allclose(np_(s[b]), os_, rtol=1e-6, atol=1e-12)
        s1, i1 = gb.problem(b).score_pool(torch.as_tensor(Xq[b]).cuda(), B200.EI(best[b]), k)
        assert torch.equal(i1, i[b]) and torch.equal(s1, s[b])


def test_single_launch_fit_equals_the_pipeline(cuda_device):
    """Config-4 shape: 100 problems of n=256, d=10 (above the default threshold: k_fit_small) against each problem
    alone through the multi-kernel pipeline (CudaGP, batch = 1): value, gradient, and the Adam trajectory of train()."""
    import bbo_b200 as B200
    from bbo_b200.batch import CudaGPBatch
    from _gpu_helpers import gp_from_params
    B, n, d = 100, 256, 10
    ps, X, y, theta = _batch(B, n, d, O.RBF, 4100)
    gb = CudaGPBatch(torch.as_tensor(X), torch.as_tensor(y)[..., None], "rbf_vmap", theta=torch.as_tensor(theta),
                     noise_variance=1e-4, epochs=10, lr=0.02)
    vals, grads = gb.value_and_grad()
    for b in (0, 7, 99):
        gp = gp_from_params(ps[b], X[b], y[b], epochs=10, lr=0.02)
        v, g_ls, g_var, g_c = gp.value_and_grad()                      # raw-parameter gradients
        sg = O.sigmoid(ps[b].raw_lengthscale).numpy()
        assert abs(float(vals[b]) - float(v)) <= 1e-10 * abs(float(v))
        np.testing.assert_allclose(np_(grads[b, :d]) * sg, np_(g_ls), rtol=1e-8, atol=1e-10 * np.abs(np_(g_ls)).max())
        assert abs(float(grads[b, d + 1]) - float(g_c)) <= 1e-8 * abs(float(g_c)) + 1e-12
    gb.train()
    for b in (0, 99):
        gp = gp_from_params(ps[b], X[b], y[b], epochs=10, lr=0.02)
        gp.train()
        np.testing.assert_allclose(np_(gb.last_values[:, b]), np_(gp.last_values), rtol=1e-9)
        got = np_(gb.theta[b])
        want = np.concatenate([[float(gp._mean_module.constant)], np_(gp._cov_module.base_kernel.lengthscale),
                               [float(gp._cov_module.variance)]])
        np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-12)
    assert int(gb.info.abs().sum()) == 0


def test_single_launch_fit_reports_non_pd(cuda_device, single_launch):
    """A duplicated observation with zero noise makes one problem's K singular: that problem's info word is set
    (the leading minor, as torch.linalg.cholesky reports it), the others are untouched."""
    from bbo_b200.batch import CudaGPBatch
    B, n, d = 8, 100, 3
    ps, X, y, theta = _batch(B, n, d, O.RBF, 77)
    X[3, 40] = X[3, 10]
    gb = CudaGPBatch(torch.as_tensor(X), torch.as_tensor(y)[..., None], "rbf_vmap", theta=torch.as_tensor(theta),
                     noise_variance=0.0, epochs=1)
    with pytest.raises(torch.linalg.LinAlgError):
        gb.value_and_grad()
