"""Appends observations to a factorised n=4096 state a few times (profiling aid for bbo_gp_append)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bbo_b200 as B
from bbo_b200.benchmarks import Ackley, make_observations
n, m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096, 8
X, y = make_observations(Ackley(20), n + 4 * m, 7)
cov = B.ScaleKernel(B.RBFKernel(ard_dim=20)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(0.0); cov.variance.fill_(0.5413)
Xt, yt = torch.as_tensor(X).cuda(), torch.as_tensor(y).reshape(-1, 1).cuda()
gp = B.CudaGP(Xt[:n], yt[:n], B.ConstantMean().double(), cov, noise_variance=1e-4)
gp.state
torch.cuda.synchronize()
for i in range(4):
    t0 = time.perf_counter()
    gp.append(Xt[n + i * m:n + (i + 1) * m], yt[n + i * m:n + (i + 1) * m])
    torch.cuda.synchronize()
    print("append %d: %.3f ms wall" % (i, (time.perf_counter() - t0) * 1e3))
