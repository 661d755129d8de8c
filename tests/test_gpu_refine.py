"""precision="tf32+refine": the TF32 tensor-core pass only short-lists; the short-list is re-scored in float64 and
the selection is made on float64 scores.  Gates: returned scores float64-exact (1e-6 vs the oracle, 1e-9 vs the
float64 mode), selected indices IDENTICAL to the float64 mode / the oracle, on the shapes of all five BASELINE
configurations including the headline shape n=4096, d=20, l=0.5, noise=1e-4."""
import numpy as np
import pytest
import torch

import bbo_b200 as B
from _golden import O
from _gpu_helpers import gp_from_params, np_, synth_problem
from bbo_b200.benchmarks import Ackley, Levy, Rastrigin, Branin2D, make_observations
from bbo_b200.pool import uniform_pool

pytestmark = pytest.mark.gpu


def _fixed(kind, d, ls, noise=1e-4, iso=False):
    raw = float(np.log(np.expm1(ls)))
    return O.GPParams(kind=kind, raw_lengthscale=(np.float64(raw) if iso else np.full(d, raw)),
                      raw_variance=float(np.log(np.expm1(1.0))), constant=0.0, noise=noise,
                      pairwise_eps=(1e-6 if kind == O.MATERN52 else 0.0))


def _same_selection(gp, pool, acq, k, must_certify=True):
    """certified  =>  the float64 selection, with float64-exact scores.  (Uncertified results are allowed to differ:
    that is what the evidence is for; precision="auto" then re-scores in float64.)"""
    s64, i64 = gp.score_pool(pool, acq, k, precision="fp64")
    sr, ir = gp.score_pool(pool, acq, k, precision="tf32+refine")
    cert = gp.refine_certificate()
    if must_certify:
        assert cert["certified"], cert
    if cert["certified"]:
        assert torch.equal(ir, i64), (ir, i64, cert)
        a, b = np_(sr), np_(s64)
        assert np.all(np.abs(a - b) <= 1e-9 * np.abs(b) + 1e-300), (a, b)
    return s64, i64, cert


def test_headline_shape_against_the_oracle(cuda_device):
    """n=4096, d=20, ARD RBF l=0.5, s2=1, noise=1e-4 (bench.py's workload): two TF32 chunks, oracle on the whole pool."""
    n, d, q, k = 4096, 20, 20000, 8
    X, y = make_observations(Ackley(d), n, 20251019)
    p = _fixed(O.RBF, d, 0.5)
    gp = gp_from_params(p, X, y)
    best = float(y.max())
    Xq = np.random.default_rng(20251020).random((q, d))
    pool = torch.as_tensor(Xq).cuda()
    sr, ir = gp.score_pool(pool, B.EI(best), k, precision="tf32+refine")
    cert = gp.refine_certificate()
    oi, os_, omu, ovar, osc = O.score_pool(p, X, y, Xq, "ei", best, k=k)
    assert list(np_(ir)) == list(oi), (np_(ir), oi, cert)
    assert np.all(np.abs(np_(sr) - os_) <= 1e-6 * np.abs(os_) + 1e-15)
    assert cert["shortlist"] == 64 and cert["certified"], cert
    # plain TF32 on the same shape, for the record: absolute posterior errors against the oracle
    _, _, mu, var, _ = gp.score_pool(pool, B.EI(best), k, precision="tf32", return_all=True)
    prior = 1.0 + 1e-4
    assert np.abs(np_(mu) - omu.numpy()).max() <= 1e-3 * max(1.0, np.abs(omu.numpy()).max())
    assert np.abs(np_(var) - ovar.numpy()).max() <= 1e-3 * prior


@pytest.mark.parametrize("name,func,n,d,q,kind,ls,iso", [
    ("config1", Branin2D(), 100, 2, 8192, O.RBF, 0.3, True),
    ("config2", Ackley(20), 1024, 20, 1 << 18, O.RBF, 0.5, False),
    ("config3", Rastrigin(50), 4096, 50, 1 << 16, O.MATERN52, 1.3, False),
    ("config4", Ackley(10), 256, 10, 1 << 16, O.RBF, 0.5, False),
    ("config5", Levy(100), 8192, 100, 1 << 16, O.RBF, 1.5, False),
    ("headline", Ackley(20), 4096, 20, 1 << 18, O.RBF, 0.5, False),
])
def test_refine_selection_equals_fp64_mode_on_config_shapes(name, func, n, d, q, kind, ls, iso, cuda_device):
    X, y = make_observations(func, n, 7)
    p = _fixed(kind, d, ls, noise=1e-4, iso=iso)
    gp = gp_from_params(p, X, y)
    best = float(y.max())
    pool = uniform_pool(11, 0, q, d)
    # config 1 (n=100, d=2, l=0.3: cond(K) ~ 1e8) is the regime where one-pass TF32 is NOT accurate (observed
    # |TF32 - float64| score errors of 0.02 (EI) and 0.09 (UCB) against k-th scores of 0.027 / 1.15): the evidence
    # refuses to certify those selections, and precision="auto" falls back to float64
    tiny = name == "config1"
    for acq, k in ((B.EI(best), 8), (B.UCB(1.8), 8), (B.PI(best), 3)):
        _, _, cert = _same_selection(gp, pool, acq, k, must_certify=not tiny)
        sa, ia = gp.score_pool(pool, acq, k, precision="auto")
        s64, i64 = gp.score_pool(pool, acq, k, precision="fp64")
        assert torch.equal(ia, i64) and torch.allclose(sa, s64, rtol=1e-9, atol=0)
        assert gp.last_auto_path == ("tf32+refine" if cert["certified"] else "fp64")
    if tiny:
        return
    # float32 candidates: the short-list rows are gathered from the float32 pool
    p32 = pool.float()
    s64, i64 = gp.score_pool(p32, B.EI(best), 8, precision="fp64")
    sr, ir = gp.score_pool(p32, B.EI(best), 8, precision="tf32+refine")
    assert torch.equal(ir, i64)
    assert torch.allclose(sr, s64, rtol=1e-9, atol=0)


def test_refine_host_pool_sharding_and_running(cuda_device):
    """The host-pool path defers the re-score to the end of the pool; shards merged through `running` and an
    index offset give the one-pass result."""
    n, d, q, k = 1024, 20, 300000, 8
    X, y = make_observations(Ackley(d), n, 3)
    p = _fixed(O.RBF, d, 0.5)
    gp = gp_from_params(p, X, y)
    acq = B.EI(float(y.max()))
    pool = uniform_pool(5, 0, q, d)
    s64, i64 = gp.score_pool(pool, acq, k, precision="fp64")
    hs, hi = gp.score_pool(pool.cpu().pin_memory(), acq, k, precision="tf32+refine", host_pool=True,
                           index_offset=1000)
    cert = gp.refine_certificate()
    assert torch.equal(hi, (i64 + 1000).cpu()) and torch.allclose(hs, s64.cpu(), rtol=1e-9, atol=0)
    assert cert["certified"], cert
    half = 150000
    r = gp.score_pool(pool[:half], acq, k, precision="tf32+refine")
    s2, i2 = gp.score_pool(pool[half:], acq, k, precision="tf32+refine", index_offset=half, running=r)
    assert torch.equal(i2, i64) and torch.allclose(s2, s64, rtol=1e-9, atol=0)


@pytest.mark.parametrize("q,k", [(50, 8), (127, 32), (2049, 1), (5000, 32), (40000, 16)])
def test_refine_small_pools_and_k(q, k, cuda_device):
    p, X, y = synth_problem(91, 700, 12, kind=O.MATERN52, noise=1e-3, ls_mean=0.0)
    gp = gp_from_params(p, X, y)
    pool = uniform_pool(q, 0, q, 12)
    _same_selection(gp, pool, B.EI(float(y.max())), k)
    _same_selection(gp, pool, B.UCB(2.5), k)
    with pytest.raises(ValueError):
        gp.score_pool(pool, B.UCB(), 33, precision="tf32+refine")


def test_refine_ties_resolve_to_the_lowest_index(cuda_device):
    """A pool far from every observation with a huge incumbent: EI underflows to exactly 0 for every candidate in
    both arithmetics; the stable order (trial.py:289-294) then selects the lowest indices."""
    p, X, y = synth_problem(5, 300, 6, kind=O.RBF, noise=1e-3, ls_mean=-2.5)
    gp = gp_from_params(p, X, y)
    pool = uniform_pool(3, 0, 30000, 6)
    s64, i64 = gp.score_pool(pool, B.EI(1e3), 8, precision="fp64")
    sr, ir = gp.score_pool(pool, B.EI(1e3), 8, precision="tf32+refine")
    assert float(s64.max()) == 0.0
    assert list(np_(i64)) == list(range(8)) and torch.equal(ir, i64) and torch.equal(sr, s64)
    # duplicated rows: equal scores, the first copy wins
    pool2 = torch.cat([pool[:5000], pool[:5000]])
    s64, i64 = gp.score_pool(pool2, B.UCB(), 8, precision="fp64")
    sr, ir = gp.score_pool(pool2, B.UCB(), 8, precision="tf32+refine")
    assert torch.equal(ir, i64)
    lo = np_(i64).reshape(4, 2)          # four best rows, each followed by its copy: lower index first
    assert np.all(lo[:, 1] == lo[:, 0] + 5000) and np.all(lo[:, 0] < 5000)


@pytest.mark.parametrize("order", ["ascending", "descending", "blocks"])
def test_selection_on_pools_sorted_by_score(order, cuda_device):
    """Worst cases of the threshold-filtered selection (k_acq_eval / k_filter_heavy / k_merge_level /
    k_merge_survivors): a pool sorted by ASCENDING score makes every candidate of every chunk beat the running list
    (all blocks take the heavy path, thousands of survivors per chunk, every chunk), descending order empties the
    filter after the short first chunk, and `blocks` mixes light and heavy blocks inside one segment.  The selection
    must be the float64 one in all three precisions, with several chunks per pool."""
    n, d, q, k = 600, 20, 60000, 8
    p, X, y = synth_problem(2024, n, d, kind=O.RBF, noise=1e-3, ls_mean=-0.3)
    gp = gp_from_params(p, X, y, q_chunk=8192)          # 8 chunks
    acq = B.UCB()
    pool = torch.as_tensor(np.random.default_rng(3).random((q, d))).cuda()
    _, _, _, _, sc = gp.score_pool(pool, acq, k, precision="fp64", return_all=True)
    idx = torch.argsort(sc)
    if order == "descending":
        idx = idx.flip(0)
    elif order == "blocks":                              # ascending runs of 300 candidates between shuffled stretches
        perm = torch.randperm(q, generator=torch.Generator().manual_seed(1)).cuda()
        idx = torch.where((torch.arange(q, device="cuda") // 300) % 2 == 0, idx, perm)
        idx = torch.unique(idx)                          # a valid (shorter) pool of distinct candidates
    pool2 = pool[idx].contiguous()
    s64, i64 = gp.score_pool(pool2, acq, k, precision="fp64")
    want = torch.argsort(sc[idx], descending=True, stable=True)[:k]
    assert torch.equal(i64, want)
    st, it = gp.score_pool(pool2, acq, k, precision="tf32")
    assert len(set(np_(it).tolist()) & set(np_(i64).tolist())) >= k - 2      # TF32-accurate scores
    sr, ir = gp.score_pool(pool2, acq, k, precision="tf32+refine")
    assert gp.refine_certificate()["certified"]
    assert torch.equal(ir, i64)
    assert np.allclose(np_(sr), np_(s64), rtol=1e-9, atol=0)
    # k = 32 (short-list of 256): the widest lists the merge kernels see
    s64b, i64b = gp.score_pool(pool2, acq, 32, precision="fp64")
    srb, irb = gp.score_pool(pool2, acq, 32, precision="tf32+refine")
    assert gp.refine_certificate()["certified"]
    assert torch.equal(irb, i64b)
    # automatic chunking: a short first chunk (one wave) followed by one long chunk
    gp2 = gp_from_params(p, X, y)
    sr2, ir2 = gp2.score_pool(pool2, acq, k, precision="tf32+refine")
    assert gp2.refine_certificate()["certified"]
    assert torch.equal(ir2, i64)
    st2, it2 = gp2.score_pool(pool2, acq, k, precision="tf32")
    assert torch.equal(it2, it) and torch.equal(st2, st)          # chunking does not change a bit
