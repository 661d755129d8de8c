"""GPU: warped kernels and learned means against the reference's own outputs
(tests/golden/warp_cases.npz) and against the CPU oracle chain."""
import numpy as np
import pytest
import torch

import bbo_b200 as B
from _golden import O
from _gpu_helpers import np_
from _warp_cases import WARP_CASES, WarpCase
from test_oracle_warp import oracle_chain

pytestmark = pytest.mark.gpu
F64 = torch.float64


def build_modules(c, trained=False):
    if c.name == "kumar":
        cov = B.ScaleKernel(B.WarpKernel(B.Matern52Kernel(ard_dim=c.d), B.KumarWarp()))
        mean = B.ConstantMean()
    else:
        cov = B.ScaleKernel(B.WarpKernel(B.RBFKernel(ard_dim=4), B.MLPWarp(c.d, [8, 4])))
        mean = B.MLPMean(5, B.MLPWarp(c.d, [5]))
    cov, mean = cov.double(), mean.double()
    cov.load_state_dict({k: torch.tensor(v) for k, v in c.group("tcov" if trained else "cov").items()})
    mean.load_state_dict({k: torch.tensor(v) for k, v in c.group("tmean" if trained else "mean").items()})
    return mean, cov


def make_gp(c, trained=False, **kw):
    mean, cov = build_modules(c, trained)
    gp = B.surrogate(torch.as_tensor(c.X), torch.as_tensor(c.Y), mean, cov, noise_variance=c.noise, **kw)
    assert isinstance(gp, B.CudaWarpedGP)
    return gp, mean, cov


@pytest.mark.parametrize("name", WARP_CASES)
def test_objective_value_and_all_gradients_match_reference_autograd(name, cuda_device):
    c = WarpCase(name)
    gp, mean, cov = make_gp(c)
    v = gp.objective()
    assert v.shape == (1, 1)
    assert abs(float(v) - float(c.value)) <= 1e-6 * abs(float(c.value))
    v.backward()
    ref = [(cov, c.group("gcov")), (mean, c.group("gmean"))]
    for mod, grads in ref:
        named = dict(mod.named_parameters())
        assert set(named) == set(grads)
        for k, g in grads.items():
            got = np_(named[k].grad)
            assert np.allclose(got, g, rtol=1e-6, atol=1e-6 * max(np.abs(g).max(), 1e-9)), (k, got, g)


@pytest.mark.parametrize("name", WARP_CASES)
@pytest.mark.parametrize("trained", [False, True])
def test_predict_matches_reference(name, trained, cuda_device):
    c = WarpCase(name)
    gp, _, _ = make_gp(c, trained)
    mu, cov = gp.predict(torch.as_tensor(c.Xq))
    rmu = c.train_mu if trained else c.mu
    rvar = c.train_var_diag if trained else c.var_diag
    assert mu.shape == (c.q, 1) and cov.shape == (c.q, c.q)
    assert np.allclose(np_(mu), rmu, rtol=1e-6, atol=1e-8)
    assert np.allclose(np_(cov.diagonal()), rvar, rtol=1e-6, atol=1e-9)
    mu3, var3 = gp.predict(torch.as_tensor(c.Xq).reshape(c.q, 1, c.d))      # the BODesigner convention
    assert mu3.shape == (c.q, 1, 1)
    assert np.allclose(np_(mu3).reshape(-1), rmu.reshape(-1), rtol=1e-6, atol=1e-8)
    assert np.allclose(np_(var3).reshape(-1), rvar, rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("name", WARP_CASES)
def test_train_follows_reference_adam(name, cuda_device):
    c = WarpCase(name)
    gp, mean, cov = make_gp(c, lr=0.01, epochs=8)
    gp.train()
    for mod, ref in ((cov, c.group("tcov")), (mean, c.group("tmean"))):
        for k, v in mod.state_dict().items():
            assert np.allclose(np_(v), ref[k], rtol=1e-5, atol=1e-6), k
    assert gp.last_values.shape == (8,)
    assert abs(float(gp.last_values[0]) - float(c.value)) <= 1e-6 * abs(float(c.value))
    mu, var = gp.predict_diag(torch.as_tensor(c.Xq))
    assert np.allclose(np_(mu), c.train_mu.reshape(-1), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", WARP_CASES)
@pytest.mark.parametrize("precision", ["fp64", "tf32"])
def test_score_pool_through_the_warp(name, precision, cuda_device):
    c = WarpCase(name)
    gp, _, _ = make_gp(c)
    _, _, (p, warp, mean_fn) = oracle_chain(c)
    rng = np.random.default_rng(3)
    Xq = rng.random((3000, c.d))
    with torch.no_grad():
        X, Y, Xqt = torch.tensor(c.X), torch.tensor(c.Y), torch.tensor(Xq)
        omu, ovar = O.predict_diag(p, warp(X), Y - mean_fn(X), warp(Xqt))
        omu = omu.reshape(-1) + mean_fn(Xqt).reshape(-1)
        best_f = float(c.Y.max())
        osc = O.ei(omu, ovar.reshape(-1).clamp_min(0).sqrt(), best_f)
    top_s, top_i, mu, var, sc = gp.score_pool(torch.as_tensor(Xq).cuda(), B.EI(best_f), 4, precision=precision,
                                              return_all=True)
    if precision == "fp64":
        assert np.allclose(np_(mu), omu.numpy(), rtol=1e-6, atol=1e-8)
        assert np.allclose(np_(var), ovar.numpy().reshape(-1), rtol=1e-6, atol=1e-10)
    else:
        # TF32 mode.  These two small low-dimensional problems are badly conditioned (d=3, noise 1e-4:
        # |L^-1| reaches 1/sqrt(noise)), which amplifies the 2^-11 operand rounding of |L^-1 k*|^2; the 1e-3
        # gate itself is asserted on well-conditioned problems in test_gpu_tf32.py.  Here the point is that
        # the warp prologue feeds the tensor-core path: 1e-2 of the prior scale.
        prior = float(O.softplus(p.raw_variance)) + p.noise
        assert np.abs(np_(mu) - omu.numpy()).max() <= 1e-2 * max(1.0, float(np.abs(c.Y).max()))
        assert np.abs(np_(var) - ovar.numpy().reshape(-1)).max() <= 1e-2 * prior
    if precision == "fp64":
        order = np.argsort(-osc.numpy(), kind="stable")[:4]
        assert np.array_equal(np_(top_i), order)
        assert np.allclose(np_(top_s), osc.numpy()[order], rtol=1e-6, atol=1e-12)
    else:
        assert float(osc[int(top_i[0])]) >= float(osc.max()) * (1 - 5e-2) - 1e-9


def test_kumar_prologue_kernel_matches_module(cuda_device):
    w = B.KumarWarp().double().cuda()
    with torch.no_grad():
        w.alpha.fill_(0.7)
        w.beta.fill_(-0.4)
    c = WarpCase("kumar")
    mean, cov = build_modules(c)
    cov.base_kernel.warp_module = w
    gp = B.CudaWarpedGP(torch.as_tensor(c.X), torch.as_tensor(c.Y), mean, cov, noise_variance=c.noise)
    x = torch.rand(1000, c.d, dtype=F64, device="cuda")
    x[0, 0], x[1, 1] = 0.0, 1.0                      # clipped to [1e-6, 1 - 1e-6]
    for dt, tol in ((F64, 1e-14), (torch.float32, 2e-6)):
        got = gp._warp_pool(x.to(dt))
        with torch.no_grad():
            want = w(x)
        assert got.dtype == dt and torch.allclose(got.to(F64), want, rtol=tol, atol=tol)


def test_plain_modules_still_take_the_fused_path(cuda_device):
    c = WarpCase("kumar")
    gp = B.surrogate(torch.as_tensor(c.X), torch.as_tensor(c.Y), B.ConstantMean().double(),
                     B.ScaleKernel(B.RBFKernel(ard_dim=c.d)).double(), noise_variance=1e-4)
    assert isinstance(gp, B.CudaGP)
    with pytest.raises(NotImplementedError):
        B.CudaGP(torch.as_tensor(c.X), torch.as_tensor(c.Y), *build_modules(c))
