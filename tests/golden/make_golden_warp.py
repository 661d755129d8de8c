"""Golden vectors for the warped-kernel / learned-mean options from the REAL reference.

Run in the build container only:  python tests/golden/make_golden_warp.py  -> tests/golden/warp_cases.npz
Two cases through the unmodified reference classes (gp/kernels.py:59-68 WarpKernel, gp/warpers.py
KumarWarp / MLPWarp, gp/means.py:28-35 MLPMean, gp/objectives.py:34-45 + autograd, gp/gp.py train/predict):
  kumar : ScaleKernel(WarpKernel(Matern52Kernel(ard), KumarWarp())) + ConstantMean
  mlp   : ScaleKernel(WarpKernel(RBFKernel(ard_dim=4), MLPWarp(3,[8,4]))) + MLPMean(5, MLPWarp(3,[5]))
Stored per case: inputs, every parameter (by state-dict key), the objective value, the autograd gradient of
every parameter, predict() before training, and parameters + predict() after a short GP.train().
"""
import copy
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, _install_gpytorch_stub  # noqa: E402


def main():
    sys.path.insert(0, REF)
    _install_gpytorch_stub()
    from bbo.algorithms.surrogates.gp import GP
    from bbo.algorithms.surrogates.gp.kernels import Matern52Kernel, RBFKernel, ScaleKernel, WarpKernel
    from bbo.algorithms.surrogates.gp.means import ConstantMean, MLPMean
    from bbo.algorithms.surrogates.gp.objectives import neg_log_marginal_likelihood
    from bbo.algorithms.surrogates.gp.warpers import KumarWarp, MLPWarp

    f64 = torch.float64
    out = {}

    def build(name):
        torch.manual_seed(4242 if name == "kumar" else 4343)
        if name == "kumar":
            n, d, q, noise = 40, 3, 30, 1e-4
            base = Matern52Kernel(ard_dim=d)
            cov = ScaleKernel(WarpKernel(base, KumarWarp()))
            mean = ConstantMean()
        else:
            n, d, q, noise = 50, 3, 30, 1e-2
            base = RBFKernel(ard_dim=4)
            cov = ScaleKernel(WarpKernel(base, MLPWarp(d, [8, 4])))
            mean = MLPMean(5, MLPWarp(d, [5]))
        cov, mean = cov.double(), mean.double()
        g = torch.Generator().manual_seed(99 if name == "kumar" else 98)
        X = torch.rand(n, d, dtype=f64, generator=g)
        w = torch.randn(d, 1, dtype=f64, generator=g)
        Y = torch.sin(3.0 * X @ w) + 0.1 * torch.randn(n, 1, dtype=f64, generator=g)
        Y = (Y - Y.mean(0)) / (Y.std(0) + 1e-6)
        Xq = torch.rand(q, d, dtype=f64, generator=g)
        with torch.no_grad():
            # the MLP case gets short lengthscales: tanh features cluster, and a badly conditioned K would
            # turn LAPACK-vs-anything rounding into 1e-8 differences that say nothing about parity
            base.lengthscale.copy_(0.2 * torch.randn(base.lengthscale.shape, dtype=f64, generator=g)
                                   - (0.3 if name == "kumar" else 1.5))
            cov.variance.copy_(torch.tensor(0.3, dtype=f64))
            if name == "kumar":
                cov.base_kernel.warp_module.alpha.copy_(torch.tensor([0.4], dtype=f64))
                cov.base_kernel.warp_module.beta.copy_(torch.tensor([-0.3], dtype=f64))
                mean.constant.copy_(torch.tensor(0.1, dtype=f64))
        return n, d, q, noise, X, Y, Xq, mean, cov

    for name in ("kumar", "mlp"):
        n, d, q, noise, X, Y, Xq, mean, cov = build(name)
        pre = name + "/"
        out[pre + "meta"] = np.array([n, d, q], dtype=np.int64)
        out[pre + "noise"] = np.array(noise)
        out[pre + "X"], out[pre + "Y"], out[pre + "Xq"] = X.numpy(), Y.numpy(), Xq.numpy()
        for k, v in mean.state_dict().items():
            out[pre + "mean/" + k] = v.numpy().copy()
        for k, v in cov.state_dict().items():
            out[pre + "cov/" + k] = v.numpy().copy()
        val = neg_log_marginal_likelihood(X, Y, mean, cov, noise)
        val.backward()
        out[pre + "value"] = val.detach().numpy().reshape(())
        for k, p in mean.named_parameters():
            out[pre + "gmean/" + k] = p.grad.numpy().copy()
        for k, p in cov.named_parameters():
            out[pre + "gcov/" + k] = p.grad.numpy().copy()
        gp = GP(X, Y, mean, cov, noise_variance=noise)
        with torch.no_grad():
            mu, var = gp.predict(Xq)
        out[pre + "mu"], out[pre + "var_diag"] = mu.numpy(), var.diagonal().numpy().copy()
        # short training run with the reference's own loop (gp.py:54-63)
        mean2, cov2 = copy.deepcopy(mean), copy.deepcopy(cov)
        for p in list(mean2.parameters()) + list(cov2.parameters()):
            p.grad = None
        gp2 = GP(X, Y, mean2, cov2, noise_variance=noise, lr=0.01, epochs=8)
        gp2.train()
        for k, v in mean2.state_dict().items():
            out[pre + "tmean/" + k] = v.numpy().copy()
        for k, v in cov2.state_dict().items():
            out[pre + "tcov/" + k] = v.numpy().copy()
        gp3 = GP(X, Y, mean2, cov2, noise_variance=noise)      # fresh cache (gp.py:66 keeps a stale one)
        with torch.no_grad():
            mu_t, var_t = gp3.predict(Xq)
        out[pre + "train_mu"], out[pre + "train_var_diag"] = mu_t.numpy(), var_t.diagonal().numpy().copy()
        print(name, "value", float(val), "keys", [k for k in out if k.startswith(pre + "gcov/")])
    np.savez_compressed(os.path.join(HERE, "warp_cases.npz"), **out)
    print("warp golden written")


if __name__ == "__main__":
    main()
