"""Generate golden vectors from the REAL reference (songlei00/bbo at /root/reference).

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
Writes tests/golden/gp_cases.npz, acq_cases.npz, select_cases.npz.

The reference ships no golden vectors for the GP path (its tests are shape /
property checks), so the pin is "outputs of the unmodified
reference on seeded inputs".  gpytorch is not installed and
bbo/algorithms/designers/__init__.py:4 imports it transitively, so a stub
module is pre-seeded in sys.modules before importing EI/UCB; no
reference arithmetic lives in the stub.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("BBO_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _install_gpytorch_stub():
    if "gpytorch" in sys.modules:
        return

    # exactly the names imported at gpytorch_gp.py:6-13 and gpytorch_gp/kernels.py:2
    wanted = {
        "gpytorch": [],
        "gpytorch.models": ["ExactGP"],
        "gpytorch.constraints": [],
        "gpytorch.constraints.constraints": ["Interval"],
        "gpytorch.distributions": ["MultivariateNormal"],
        "gpytorch.means": ["ConstantMean"],
        "gpytorch.kernels": ["RBFKernel", "MaternKernel", "ScaleKernel", "Kernel"],
        "gpytorch.likelihoods": ["GaussianLikelihood"],
        "gpytorch.mlls": ["ExactMarginalLogLikelihood"],
        "gpytorch.priors": ["torch_priors"],
    }
    for name, attrs in wanted.items():
        m = types.ModuleType(name)
        m.__path__ = []                      # importable as a package
        for a in attrs:
            setattr(m, a, type(a, (torch.nn.Module,), {}))
        if name == "gpytorch.priors":        # module-level calls at gpytorch_gp.py:22-26
            tp = types.SimpleNamespace(GammaPrior=lambda *a, **k: None,
                                       LogNormalPrior=lambda *a, **k: None)
            m.torch_priors = tp
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)


def main():
    sys.path.insert(0, REF)
    _install_gpytorch_stub()
    from bbo.algorithms.surrogates.gp import GP
    from bbo.algorithms.surrogates.gp import kernel_impl
    from bbo.algorithms.surrogates.gp.kernels import (RBFKernel, Matern52Kernel, ScaleKernel,
                                                        TemplateKernel)
    from bbo.algorithms.surrogates.gp.means import ConstantMean
    from bbo.algorithms.surrogates.gp.objectives import neg_log_marginal_likelihood
    from bbo.algorithms.designers.bo.acquisitions import EI, UCB
    from bbo.shared.trial import Trial, Measurement, get_best_trials
    from bbo.shared.base_study_config import ProblemStatement, MetricInformation, ObjectiveMetricGoal

    f64 = torch.float64
    cases = [
        # name, kind, impl, ard, scale, n, d, q, noise, eps
        ("rbf_ard_scale", "rbf", "vmap", True, True, 48, 3, 40, 1e-4, 0.0),
        ("m52_iso_default", "matern52", "vmap", False, False, 40, 2, 33, 1e-6, 1e-6),
        ("m52_ard_scale", "matern52", "vmap", True, True, 64, 5, 50, 1e-4, 1e-6),
        ("rbf_iso_scale", "rbf", "vmap", False, True, 30, 2, 20, 1e-3, 0.0),
        ("m52_cdist_ard", "matern52", "cdist", True, True, 36, 4, 25, 1e-4, 0.0),
        ("rbf_cdist_ard", "rbf", "cdist", True, False, 36, 4, 25, 1e-4, 0.0),
        ("m52_ard_scale_d20", "matern52", "vmap", True, True, 96, 20, 64, 1e-6, 1e-6),
        ("rbf_ard_scale_d20", "rbf", "vmap", True, True, 130, 20, 70, 1e-6, 0.0),
    ]
    out = {}
    names = []
    for ci, (name, kind, impl, ard, scale, n, d, q, noise, eps) in enumerate(cases):
        g = torch.Generator().manual_seed(1000 + ci)
        X = torch.rand(n, d, dtype=f64, generator=g)
        w = torch.randn(d, 1, dtype=f64, generator=g)
        Y = torch.sin(3.0 * X @ w) + 0.1 * torch.randn(n, 1, dtype=f64, generator=g)
        Y = (Y - Y.mean(0)) / (Y.std(0) + 1e-6)
        Xq = torch.rand(q, d, dtype=f64, generator=g)
        raw_ls = (0.3 * torch.randn(d if ard else (), dtype=f64, generator=g) - 0.4) if ard \
            else torch.tensor(-0.3, dtype=f64)
        raw_var = torch.tensor(0.4, dtype=f64)
        const = torch.tensor(0.15, dtype=f64)

        fn = getattr(kernel_impl, f"{kind}_{impl}")
        base = TemplateKernel(fn, ard_dim=d if ard else None).double()
        cov = ScaleKernel(base).double() if scale else base
        mean = ConstantMean().double()
        with torch.no_grad():
            base.lengthscale.copy_(raw_ls)
            if scale:
                cov.variance.copy_(raw_var)
            mean.constant.copy_(const)

        # A1-A3: covariance
        with torch.no_grad():
            Kxx = cov(X, X)
            Kqx = cov(Xq, X)
        # A6: value + autograd gradient
        val = neg_log_marginal_likelihood(X, Y, mean, cov, noise)
        val.backward()
        g_ls = base.lengthscale.grad.clone()
        g_var = cov.variance.grad.clone() if scale else torch.tensor(float("nan"), dtype=f64)
        g_c = mean.constant.grad.clone()
        # A8: posterior (2-D path only; the 3-D path is broken in the reference)
        gp = GP(X, Y, mean, cov, noise_variance=noise)
        with torch.no_grad():
            mu, var = gp.predict(Xq)
            chol, kinvy = gp._cached_chol, gp._cached_kinvy
        # A7: a short Adam run with the reference's own train()
        base2 = TemplateKernel(fn, ard_dim=d if ard else None).double()
        cov2 = ScaleKernel(base2).double() if scale else base2
        mean2 = ConstantMean().double()
        with torch.no_grad():
            base2.lengthscale.copy_(raw_ls)
            if scale:
                cov2.variance.copy_(raw_var)
            mean2.constant.copy_(const)
        gp2 = GP(X, Y, mean2, cov2, noise_variance=noise, lr=0.01, epochs=12)
        gp2.train()
        with torch.no_grad():
            mu_t, var_t = gp2.predict(Xq)

        names.append(name)
        pre = name + "/"
        out[pre + "meta"] = np.array([n, d, q, int(ard), int(scale), 0 if kind == "rbf" else 1], dtype=np.int64)
        out[pre + "noise_eps"] = np.array([noise, eps], dtype=np.float64)
        out[pre + "X"] = X.numpy()
        out[pre + "Y"] = Y.numpy()
        out[pre + "Xq"] = Xq.numpy()
        out[pre + "raw_ls"] = raw_ls.numpy()
        out[pre + "raw_var"] = raw_var.numpy()
        out[pre + "const"] = const.numpy()
        out[pre + "Kxx"] = Kxx.numpy()
        out[pre + "Kqx"] = Kqx.numpy()
        out[pre + "value"] = val.detach().numpy().reshape(())
        out[pre + "g_ls"] = g_ls.numpy()
        out[pre + "g_var"] = g_var.numpy()
        out[pre + "g_c"] = g_c.numpy()
        out[pre + "mu"] = mu.numpy()
        out[pre + "var_diag"] = var.diagonal().numpy().copy()
        out[pre + "var_full"] = var.numpy()
        out[pre + "chol_diag"] = chol.diagonal().numpy().copy()
        out[pre + "kinvy"] = kinvy.numpy()
        out[pre + "train_raw_ls"] = base2.lengthscale.detach().numpy()
        out[pre + "train_raw_var"] = (cov2.variance.detach().numpy() if scale else np.array(np.nan))
        out[pre + "train_const"] = mean2.constant.detach().numpy()
        out[pre + "train_mu"] = mu_t.numpy()
        out[pre + "train_var_diag"] = var_t.diagonal().numpy().copy()
        print(f"{name}: value={float(val):.6f} |g_ls|={float(g_ls.abs().max()):.3e}")
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "gp_cases.npz"), **out)

    # ---- A9: EI / UCB through the reference classes
    g = torch.Generator().manual_seed(77)
    mu = torch.randn(256, dtype=f64, generator=g)
    sd = torch.rand(256, dtype=f64, generator=g) * 0.7 + 1e-3
    # deep tails and sigma == 0
    mu[:8] = torch.tensor([-3.0, -5.0, -8.0, -9.0, 2.0, 4.0, 0.5, 0.7], dtype=f64)
    sd[:8] = torch.tensor([0.3, 0.5, 0.9, 1.0, 0.2, 0.1, 0.0, 1e-9], dtype=f64)
    best_f = 0.6
    acq = {
        "mu": mu.numpy(), "sd": sd.numpy(), "best_f": np.array(best_f),
        "ei": EI(best_f)(mu, sd).numpy(),
        "ei_eps0": EI(best_f, 0.0)(mu, sd).numpy(),
        "ucb": UCB()(mu, sd).numpy(),
        "ucb_b3": UCB(3.0)(mu, sd).numpy(),
        "ei_f32": EI(best_f)(mu.float(), sd.float()).numpy(),
    }
    np.savez_compressed(os.path.join(HERE, "acq_cases.npz"), **acq)

    # ---- A11: selection order through get_best_trials (stable, ties -> lowest index)
    ps = ProblemStatement()
    ps.search_space.add_float_param("x0", 0.0, 1.0)
    ps.metric_information.append(MetricInformation("acquisition", ObjectiveMetricGoal.MAXIMIZE))
    rng = np.random.default_rng(5)
    scores = np.round(rng.standard_normal(200), 1)          # many ties on purpose
    trials = []
    for i, s in enumerate(scores):
        t = Trial({"x0": float(i) / 200.0})
        t.complete(Measurement({"acquisition": float(s)}))
        trials.append(t)
    sel = {"scores": scores}
    for k in (1, 5, 16, 64):
        best = get_best_trials(trials, ps, count=k)
        sel[f"top{k}"] = np.array([int(round(t.parameters["x0"].value * 200.0)) for t in best], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "select_cases.npz"), **sel)
    print("golden written")


if __name__ == "__main__":
    main()
