"""TF32-mode accuracy against the FP64 mode of the same library (development aid): max / mean signed error of mu and
var for a few shapes.  Run twice (default, BBO_KSTAR_NO_UMMA=1) to compare K* producers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B


def problem(seed, n, d, kind, ls_mean):
    rng = np.random.default_rng(seed)
    X = rng.random((n, d))
    w = rng.standard_normal(d)
    y = np.sin(2.0 * X @ w) + 0.1 * rng.standard_normal(n)
    y = (y - y.mean()) / (y.std(ddof=1) + 1e-6)
    base = (B.RBFKernel if kind == "rbf" else B.Matern52Kernel)(ard_dim=d)
    cov = B.ScaleKernel(base).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.copy_(torch.as_tensor(rng.normal(ls_mean, 0.15, d)))
        cov.variance.fill_(0.2)
    mean = B.ConstantMean().double()
    with torch.no_grad():
        mean.constant.fill_(0.05)
    return B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), mean, cov, noise_variance=1e-4), y


for n, d, kind, ls in [(254, 12, "rbf", -0.2), (1024, 20, "rbf", 0.2), (2100, 20, "rbf", 0.2),
                       (600, 24, "matern52", 0.2), (200, 4, "rbf", -2.0), (4096, 20, "rbf", 0.2)]:
    gp, y = problem(1000 + n, n, d, kind, ls)
    Xq = torch.as_tensor(np.random.default_rng(n).random((4096, d))).cuda()
    acq = B.EI(float(y.max()))
    _, _, mu64, var64, s64 = gp.score_pool(Xq, acq, 4, precision="fp64", return_all=True)
    _, _, mu32, var32, s32 = gp.score_pool(Xq, acq, 4, precision="tf32", return_all=True)
    em, ev, es = (mu32 - mu64).cpu().numpy(), (var32 - var64).cpu().numpy(), (s32 - s64).cpu().numpy()
    print(json.dumps({"n": n, "d": d, "kind": kind, "mu_max": float(np.abs(em).max()), "mu_mean": float(em.mean()),
                      "var_max": float(np.abs(ev).max()), "var_mean": float(ev.mean()),
                      "score_max": float(np.abs(es).max()), "score_scale": float(s64.abs().max())}), flush=True)
