"""GPU: the BASELINE.json configurations at their REAL sizes.

Where the CPU oracle finishes in seconds (one factorisation at n <= 8192, a few hundred candidates) the CUDA
path is compared with it directly; the pool-scale part is covered by size-independent properties (sharding,
FP64-vs-TF32 agreement of the selection, variance bounds)."""
import numpy as np
import pytest
import torch

import bbo_b200 as B
from _golden import O
from _gpu_helpers import gp_from_params, np_
from bbo_b200.benchmarks import Ackley, Levy, Rastrigin, make_observations
from bbo_b200.pool import uniform_pool

pytestmark = pytest.mark.gpu


def _params(kind, d, raw_ls, noise=1e-4):
    return O.GPParams(kind=kind, raw_lengthscale=np.full(d, raw_ls), raw_variance=0.5413, constant=0.0, noise=noise,
                      pairwise_eps=(1e-6 if kind == O.MATERN52 else 0.0))


def test_config2_ackley20_n1024_pool_2e20(cuda_device):
    """Config 2: Ackley-20D, n=1024, 2^20-candidate pool, EI: oracle on a 2048-candidate slice, sharding and
    precision-mode properties on the full pool."""
    d, n = 20, 1024
    X, y = make_observations(Ackley(d), n, 7)
    p = _params(O.RBF, d, 0.0)
    gp = gp_from_params(p, X, y)
    best = float(y.max())
    pool = uniform_pool(2, 0, 1 << 20, d)                       # float64, on the device
    s64, i64 = gp.score_pool(pool, B.EI(best), 8)
    # (a) oracle on the slice that contains the winner
    lo = int(i64[0]) // 2048 * 2048
    sl = pool[lo:lo + 2048]
    oi, os_, omu, ovar, osc = O.score_pool(p, X, y, np_(sl), "ei", best, k=1)
    assert lo + int(oi[0]) == int(i64[0])
    assert abs(float(os_[0]) - float(s64[0])) <= 1e-6 * abs(float(os_[0])) + 1e-15
    # (b) two shards merged == one pass
    half = 1 << 19
    r = gp.score_pool(pool[:half], B.EI(best), 8)
    s2, i2 = gp.score_pool(pool[half:], B.EI(best), 8, index_offset=half, running=r)
    assert torch.equal(i2, i64) and torch.equal(s2, s64)
    # (c) the TF32 mode picks a winner whose FP64 score is within 1e-3 of the FP64 optimum
    s32, i32 = gp.score_pool(pool, B.EI(best), 8, precision="tf32")
    mu, var = gp.predict_diag(pool[i32[:1]])
    ref = float(O.ei(mu.cpu(), var.sqrt().cpu(), best))
    assert ref >= float(s64[0]) * (1 - 1e-3) - 1e-12


def test_config3_rastrigin50_matern_n4096_fit_and_restarts(cuda_device):
    """Config 3: Rastrigin-50D, ARD Matern-5/2 (eps form), n=4096: FP64 objective value and gradient against the
    oracle at full size, then 64 restarts x 4096 candidates merged into one selection."""
    d, n = 50, 4096
    X, y = make_observations(Rastrigin(d), n, 7)
    p = _params(O.MATERN52, d, 1.0)
    gp = gp_from_params(p, X, y)
    v, g_ls, g_var, g_c = gp.value_and_grad()
    ov, og_ls, og_var, og_c = O.objective_value_and_grad(p, X, y)
    assert abs(float(v) - ov) <= 1e-6 * abs(ov)
    assert np.max(np.abs(np_(g_ls) - og_ls)) <= 1e-6 * np.max(np.abs(og_ls))
    assert abs(float(g_var) - og_var) <= 1e-6 * abs(og_var)
    assert abs(float(g_c) - og_c) <= 1e-6 * max(abs(og_c), 1.0)
    best = float(y.max())
    run, pools = None, []
    for s in range(64):
        pl = uniform_pool(100 + s, 0, 4096, d)
        pools.append(pl)
        run = gp.score_pool(pl, B.EI(best), 4, index_offset=s * 4096, running=run)
    allp = torch.cat(pools)
    s_all, i_all = gp.score_pool(allp, B.EI(best), 4)
    assert torch.equal(run[1], i_all) and torch.equal(run[0], s_all)
    # the winner's posterior against the oracle (one factorisation, 4 candidates)
    omu, ovar = O.predict_diag(p, X, y, np_(allp[i_all]))
    mu, var = gp.predict_diag(allp[i_all])
    assert np.allclose(np_(mu), omu.numpy().reshape(-1), rtol=1e-6, atol=1e-8)
    assert np.allclose(np_(var), ovar.numpy().reshape(-1), rtol=1e-6, atol=1e-9)


def test_config5_levy100_n8192_shard(cuda_device):
    """Config 5: Levy-100D, n=8192: objective value against the oracle, posterior of 256 candidates against the
    oracle, and the float32-candidate TF32 shard against the FP64 mode."""
    d, n = 100, 8192
    X, y = make_observations(Levy(d), n, 7)
    p = _params(O.RBF, d, 1.5)
    gp = gp_from_params(p, X, y)
    v, g_ls, g_var, g_c = gp.value_and_grad()
    ov = O.objective_value(p, X, y)
    assert abs(float(v) - ov) <= 1e-6 * abs(ov)
    assert torch.isfinite(g_ls).all() and torch.isfinite(g_var)
    pool32 = uniform_pool(5, 0, 1 << 16, d, dtype=torch.float32)
    sub = pool32[:256].double()
    omu, ovar = O.predict_diag(p, X, y, np_(sub))
    mu, var = gp.predict_diag(sub)
    assert np.allclose(np_(mu), omu.numpy().reshape(-1), rtol=1e-6, atol=1e-8)
    assert np.allclose(np_(var), ovar.numpy().reshape(-1), rtol=1e-6, atol=1e-9)
    best = float(y.max())
    s32, i32, mu32, var32, _ = gp.score_pool(pool32, B.UCB(), 8, precision="tf32", return_all=True)
    s64, i64, mu64, var64, _ = gp.score_pool(pool32, B.UCB(), 8, precision="fp64", return_all=True)
    prior = float(O.softplus(p.raw_variance)) + p.noise
    assert float((mu32 - mu64).abs().max()) <= 1e-3 * max(1.0, float(mu64.abs().max()))
    assert float((var32 - var64).abs().max()) <= 1e-3 * prior
    assert len(set(np_(i32).tolist()) & set(np_(i64).tolist())) >= 6
