"""GPU parity: fit path (covariance, objective value + gradient, Adam train) vs golden vectors
from the real reference and vs the CPU oracle.  All calls go through the C ABI."""
import numpy as np
import pytest
import torch

from _golden import CASE_NAMES, Case, O
from _gpu_helpers import gp_from_case, gp_from_params, make_modules, np_, synth_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", CASE_NAMES)
def test_covariance_golden(name, cuda_device):
    import bbo_b200 as B
    c = Case(name)
    _, cov = make_modules(c.kind, c.eps, c.raw_ls, float(c.raw_var) if c.scale else None, float(c.const))
    X, Xq = torch.as_tensor(c.X).cuda(), torch.as_tensor(c.Xq).cuda()
    np.testing.assert_allclose(np_(B.cov_matrix(cov, X, X)), c.Kxx, rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(np_(cov(Xq, X)), c.Kqx, rtol=1e-12, atol=1e-15)
    # 3-D batched form (kernel_impl.py:106-128): 2-D operand is broadcast
    K3 = np_(cov(Xq.reshape(1, *Xq.shape).expand(2, -1, -1), X))
    np.testing.assert_allclose(K3[1], c.Kqx, rtol=1e-12, atol=1e-15)


@pytest.mark.parametrize("name", CASE_NAMES)
def test_value_and_grad_golden(name, cuda_device):
    c = Case(name)
    gp = gp_from_case(c)
    val, g_ls, g_var, g_c = gp.value_and_grad()
    tol = 1e-9 if name.endswith("d20") else 1e-6     # north-star gate: 1e-6 relative
    assert abs(float(val) - float(c.value)) <= tol * abs(float(c.value))
    np.testing.assert_allclose(np_(g_ls), c.g_ls, rtol=0, atol=tol * np.abs(c.g_ls).max())
    if c.scale:
        assert abs(float(g_var) - float(c.g_var)) <= tol * abs(float(c.g_var))
    assert abs(float(g_c) - float(c.g_c)) <= tol * abs(float(c.g_c))


@pytest.mark.parametrize("name", CASE_NAMES)
def test_train_golden(name, cuda_device):
    """12 Adam epochs of GP.train (gp.py:54-63), minimising +log p like the reference (F4)."""
    c = Case(name)
    gp = gp_from_case(c, lr=0.01, epochs=12)
    gp.train()
    th = gp._theta
    np.testing.assert_allclose(np_(th.raw_ls), c.train_raw_ls, rtol=1e-6, atol=1e-8)
    if c.scale:
        assert abs(float(th.raw_var) - float(c.train_raw_var)) < 1e-7
    assert abs(float(th.const) - float(c.train_const)) < 1e-7
    assert abs(float(gp.last_values[0]) - float(c.value)) <= 1e-6 * abs(float(c.value))
    mu, var = gp.predict_diag(torch.as_tensor(c.Xq).cuda())
    np.testing.assert_allclose(np_(mu), c.train_mu[:, 0], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(np_(var), c.train_var_diag, rtol=1e-5, atol=1e-11)


@pytest.mark.parametrize("n,d,kind", [(1, 1, O.RBF), (2, 3, O.MATERN52), (63, 2, O.RBF), (64, 4, O.MATERN52),
                                      (65, 7, O.RBF), (129, 33, O.MATERN52), (300, 6, O.RBF),
                                      (515, 20, O.MATERN52)])
def test_value_and_grad_oracle_shapes(n, d, kind, cuda_device):
    """Ragged sizes around the 64/128 tile boundaries and the 32-wide feature chunk."""
    p, X, y = synth_problem(100 + n, n, d, kind=kind)
    gp = gp_from_params(p, X, y)
    val, g_ls, g_var, g_c = gp.value_and_grad()
    oval, og_ls, og_var, og_c = O.objective_value_and_grad(p, X, y)
    assert abs(float(val) - oval) <= 1e-7 * abs(oval) + 1e-9
    np.testing.assert_allclose(np_(g_ls), og_ls, rtol=0, atol=1e-6 * max(np.abs(og_ls).max(), 1e-3))
    assert abs(float(g_var) - og_var) <= 1e-6 * abs(og_var) + 1e-9
    assert abs(float(g_c) - og_c) <= 1e-6 * abs(og_c) + 1e-9


def test_isotropic_no_scale(cuda_device):
    p, X, y = synth_problem(7, 150, 5, kind=O.MATERN52, ard=False, scale=False)
    gp = gp_from_params(p, X, y)
    val, g_ls, g_var, g_c = gp.value_and_grad()
    oval, og_ls, og_var, og_c = O.objective_value_and_grad(p, X, y)
    assert g_var is None and og_var is None
    assert abs(float(val) - oval) <= 1e-7 * abs(oval)
    assert abs(float(g_ls) - float(og_ls)) <= 1e-6 * abs(float(og_ls))


def test_train_sign_minus_improves_likelihood(cuda_device):
    """sign=-1 minimises the true NLL: +log p must go up; sign=+1 (reference, F4) drives it down."""
    p, X, y = synth_problem(11, 200, 4, kind=O.RBF, noise=1e-2)
    gp = gp_from_params(p, X, y, epochs=30, lr=0.05, sign=-1.0)
    gp.train()
    v = np_(gp.last_values)
    assert v[-1] > v[0]
    gp2 = gp_from_params(p, X, y, epochs=30, lr=0.05, sign=+1.0)
    gp2.train()
    v2 = np_(gp2.last_values)
    assert v2[-1] < v2[0]
    # oracle trajectory agrees
    final, ovals = O.adam_train(p, X, y, lr=0.05, epochs=30, sign=-1.0)
    np.testing.assert_allclose(v, ovals, rtol=1e-6)
    np.testing.assert_allclose(np_(gp._theta.raw_ls), np.asarray(final.raw_lengthscale), rtol=1e-6, atol=1e-8)


def test_not_positive_definite_raises(cuda_device):
    """Duplicate rows with zero noise: torch.linalg.cholesky raises at objectives.py:29."""
    X = np.tile(np.linspace(0, 1, 8).reshape(-1, 1), (2, 2))
    y = np.arange(16, dtype=np.float64)
    p = O.GPParams(kind=O.RBF, raw_lengthscale=np.float64(0.5), raw_variance=None, constant=0.0, noise=0.0)
    gp = gp_from_params(p, X, y)
    with pytest.raises(torch.linalg.LinAlgError):
        gp.value_and_grad()
    with pytest.raises(torch.linalg.LinAlgError):
        gp.predict_diag(torch.rand(3, 2, dtype=torch.float64))


def test_constructor_validation(cuda_device):
    import bbo_b200 as B
    with pytest.raises(TypeError):
        B.CudaGP(np.zeros((3, 2)), torch.zeros(3, 1))
    with pytest.raises(ValueError):
        B.CudaGP(torch.zeros(3), torch.zeros(3, 1))
