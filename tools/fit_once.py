"""Runs a few fit iterations (value + gradient) at n, d for profiling: python tools/fit_once.py [n] [d] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B
from bbo_b200.benchmarks import Ackley, make_observations
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
X, y = make_observations(Ackley(d), n, 1)
cov = B.ScaleKernel(B.RBFKernel(ard_dim=d)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(0.0); cov.variance.fill_(0.5)
gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov, noise_variance=1e-4)
for _ in range(iters):
    v, *_ = gp.value_and_grad()
torch.cuda.synchronize()
print("value", float(v))
