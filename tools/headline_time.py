"""Direct timing of CudaGP.score_pool on the headline shape (or n = argv[1]); BBO_TF32_TIMELINE=1 prints the
per-chunk K* / panel / selection timeline of every call on stderr."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B
from bbo_b200.benchmarks import Ackley, make_observations
from bbo_b200.pool import uniform_pool
D = int(sys.argv[2]) if len(sys.argv) > 2 else 20
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
X, y = make_observations(Ackley(D), n, 20251019)
cov = B.ScaleKernel(B.RBFKernel(ard_dim=D)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(float(np.log(np.expm1(0.5))))
    cov.variance.fill_(float(np.log(np.expm1(1.0))))
gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov, noise_variance=1e-4)
acq = B.EI(float(y.max()))
pool = uniform_pool(3, 0, (1 << 20) if n <= 4096 else (1 << 19), D)
Q = pool.shape[0]
for prec in ("tf32+refine", "tf32"):
    for _ in range(3):
        gp.score_pool(pool, acq, 8, precision=prec)
    torch.cuda.synchronize()
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            gp.score_pool(pool, acq, 8, precision=prec)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        print(prec, "ms/step", ms, "Mcand/s", Q / ms / 1e3, flush=True)
