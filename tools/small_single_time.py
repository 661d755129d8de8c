"""One small problem (n=256, d=10) alone: fit iteration and scoring latency (development aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B
from bbo_b200.benchmarks import Branin2D, Ackley, make_observations
def run(func, n, epochs=100):
    X, y = make_observations(func, n, 7)
    cov = B.ScaleKernel(B.RBFKernel(ard_dim=func.dim)).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(0.0); cov.variance.fill_(0.5413)
    gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov, noise_variance=1e-4, epochs=epochs, sign=-1.0, lr=0.05)
    gp.train(); torch.cuda.synchronize()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(0.0); cov.variance.fill_(0.5413)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gp.train(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / epochs, float(gp.last_values[-1])
out = {"BBO_FIT_SMALL": os.environ.get("BBO_FIT_SMALL", "default")}
for name, func, n in [("branin n=100", Branin2D(), 100), ("ackley10 n=256", Ackley(10), 256), ("ackley20 n=500", Ackley(20), 500)]:
    ms, v = run(func, n)
    out[name] = {"ms_per_epoch": ms, "last_value": v}
print(json.dumps(out))
