"""GPU: edge cases of the boundary -- empty / ragged pools, non-contiguous and float32 inputs, k at its
limits, NaN handling, workspace errors straight through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch

from _golden import O
from _gpu_helpers import gp_from_params, np_, synth_problem

pytestmark = pytest.mark.gpu


def test_empty_and_tiny_pools(cuda_device):
    import bbo_b200 as B
    p, X, y = synth_problem(1, 90, 3)
    gp = gp_from_params(p, X, y)
    s, i = gp.score_pool(torch.empty(0, 3, dtype=torch.float64, device="cuda"), B.UCB(), 4)
    assert np.isneginf(np_(s)).all() and (np_(i) == -1).all()
    for prec in ("fp64", "tf32"):
        s, i = gp.score_pool(torch.rand(1, 3, dtype=torch.float64, device="cuda"), B.UCB(), 1, precision=prec)
        assert int(i[0]) == 0 and np.isfinite(float(s[0]))
    mu, var = gp.predict_diag(torch.empty(0, 3, dtype=torch.float64, device="cuda"))
    assert mu.numel() == 0 and var.numel() == 0


def test_noncontiguous_float32_and_cpu_inputs(cuda_device):
    import bbo_b200 as B
    p, X, y = synth_problem(2, 130, 4, kind=O.MATERN52)
    # float32, non-contiguous training data (a strided view): the shim copies to contiguous float64
    Xw = torch.as_tensor(np.concatenate([X, X], axis=1)).float()[:, ::2][:, :4]
    Xw = torch.as_tensor(X).float().t().contiguous().t()          # column-major view
    assert not Xw.is_contiguous()
    mean, cov = __import__("_gpu_helpers").make_modules(p.kind, p.pairwise_eps, p.raw_lengthscale, p.raw_variance,
                                                         p.constant)
    gp = B.CudaGP(Xw, torch.as_tensor(y).float().reshape(-1, 1), mean, cov, noise_variance=p.noise)
    Xq = torch.rand(77, 8, dtype=torch.float64)[:, ::2]            # non-contiguous CPU query
    mu, var = gp.predict_diag(Xq)
    omu, ovar = O.predict_diag(p, Xw.double().numpy(), torch.as_tensor(y).float().double().numpy(), Xq.numpy())
    np.testing.assert_allclose(np_(mu), omu.numpy(), rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(np_(var), ovar.numpy(), rtol=1e-6, atol=1e-10)
    m3, v3 = gp.predict(Xq.float().unsqueeze(1))                    # float32 (q,1,d) -> float32 (q,1,1)
    assert m3.dtype == torch.float32 and m3.shape == (77, 1, 1)


def test_topk_limits_and_bad_arguments(cuda_device):
    import bbo_b200 as B
    from bbo_b200 import _lib
    p, X, y = synth_problem(3, 70, 2)
    gp = gp_from_params(p, X, y)
    Xq = torch.rand(500, 2, dtype=torch.float64, device="cuda")
    s, i, mu, var, sc = gp.score_pool(Xq, B.UCB(), _lib.TOPK_MAX, return_all=True)
    order = np.lexsort((np.arange(500), -np_(sc)))[:_lib.TOPK_MAX]
    np.testing.assert_array_equal(np_(i), order)
    with pytest.raises(ValueError):
        gp.score_pool(Xq, B.UCB(), 0)
    with pytest.raises(ValueError):
        gp.score_pool(Xq, B.UCB(), _lib.TOPK_MAX + 1)
    with pytest.raises(ValueError):
        gp.score_pool(torch.rand(5, 3, dtype=torch.float64, device="cuda"), B.UCB(), 1)
    # straight through the C ABI: a workspace that is too small is refused, nothing is launched
    lib = _lib.load()
    st = gp.state
    top_s = torch.empty(1, dtype=torch.float64, device="cuda")
    top_i = torch.empty(1, dtype=torch.int64, device="cuda")
    a = B.UCB().descriptor()
    ws = torch.empty(1024, dtype=torch.uint8, device="cuda")
    before = lib.bbo_launch_count()
    rc = lib.bbo_gp_score_pool(C.byref(st["st"]), C.byref(a), 0, _lib.ptr(Xq), 0, 500, 0, 1, _lib.ptr(top_s),
                               _lib.ptr(top_i), None, None, None, None, _lib.ptr(ws), ws.numel(), 512, 0,
                               _lib.stream_ptr())
    assert rc == 3 and lib.bbo_launch_count() == before            # BBO_ERR_WORKSPACE
    rc = lib.bbo_gp_score_pool(C.byref(st["st"]), C.byref(a), 7, _lib.ptr(Xq), 0, 500, 0, 1, _lib.ptr(top_s),
                               _lib.ptr(top_i), None, None, None, None, _lib.ptr(ws), ws.numel(), 512, 0,
                               _lib.stream_ptr())
    assert rc == 2                                                  # BBO_ERR_BAD_ARG: unknown precision


def test_nan_candidates_are_counted_and_never_selected(cuda_device):
    import bbo_b200 as B
    p, X, y = synth_problem(4, 80, 3)
    gp = gp_from_params(p, X, y)
    Xq = torch.rand(300, 3, dtype=torch.float64, device="cuda")
    Xq[[5, 17, 250]] = float("nan")
    s, i, mu, var, sc = gp.score_pool(Xq, B.EI(float(y.max())), 10, return_all=True)
    assert int(gp.last_nan_count) == 3
    assert not set(np_(i).tolist()) & {5, 17, 250}
    assert np.isnan(np_(sc)[[5, 17, 250]]).all() and np.isfinite(np.delete(np_(sc), [5, 17, 250])).all()


def test_second_device_index_is_respected(cuda_device):
    import bbo_b200 as B
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    p, X, y = synth_problem(5, 60, 2)
    gp = gp_from_params(p, X, y, device="cuda:1")
    mu, var = gp.predict_diag(torch.rand(10, 2, dtype=torch.float64))
    assert mu.device.index == 1
