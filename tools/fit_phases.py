"""Per-phase device time of one fit iteration (value + gradient): python tools/fit_phases.py [n] [d] [iters]

Phases are the library's own CUDA-event regions (bbo_profile_*): chol = covariance build + Cholesky,
trtri = triangular inverse + solves, grad = LAUUM + gradient tiles.  Also checks the value and the
gradient against a torch float64 computation on the same device (cuSOLVER Cholesky + autograd-free
closed form is not needed: value only)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bbo_b200 as B
from bbo_b200 import _lib
from bbo_b200.benchmarks import Ackley, make_observations

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 20
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
X, y = make_observations(Ackley(d), n, 1)
cov = B.ScaleKernel(B.RBFKernel(ard_dim=d)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(0.0)
    cov.variance.fill_(0.5)
gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
              noise_variance=1e-4)
lib = _lib.load()
for _ in range(3):
    v, *_ = gp.value_and_grad()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    v, g_ls, g_var, g_c = gp.value_and_grad()
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / iters
lib.bbo_profile_enable(1)
for t in ("chol", "trtri", "grad"):
    _lib.profile_read(t)
for _ in range(iters):
    gp.value_and_grad()
torch.cuda.synchronize()
ph = {}
for t in ("chol", "trtri", "grad"):
    r, ms = _lib.profile_read(t)
    ph[t] = ms / max(r, 1)
lib.bbo_profile_enable(0)

# one Adam epoch (value + gradient + update) inside GP.train: 1 eager epoch + CUDA-graph replays
E = int(os.environ.get("FIT_EPOCHS", "40"))
gp2 = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
               noise_variance=1e-4, epochs=E)
gp2.train()
torch.cuda.synchronize()
e0.record()
gp2.train()
e1.record()
torch.cuda.synchronize()
train_ms = e0.elapsed_time(e1) / E
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(0.0)
    cov.variance.fill_(0.5)

# reference value on the same device (torch float64): +log p(y)
Xd = torch.as_tensor(X, device="cuda", dtype=torch.float64)
yd = torch.as_tensor(y, device="cuda", dtype=torch.float64).reshape(-1)
ls = torch.nn.functional.softplus(torch.zeros(d, dtype=torch.float64, device="cuda"))
s2 = torch.nn.functional.softplus(torch.tensor(0.5, dtype=torch.float64, device="cuda"))
Z = Xd / ls
D2 = (Z[:, None, :] - Z[None, :, :]).pow(2).sum(-1) if n <= 2048 else torch.cdist(Z, Z).pow(2)
K = s2 * torch.exp(-0.5 * D2) + 1e-4 * torch.eye(n, dtype=torch.float64, device="cuda")
Lc = torch.linalg.cholesky(K)
al = torch.cholesky_solve(yd[:, None], Lc)[:, 0]
ref = float(-0.5 * (yd * al).sum() - torch.log(torch.diagonal(Lc)).sum() - 0.5 * n * np.log(2 * np.pi))
print(json.dumps({"n": n, "d": d, "ms_per_iter": total, "train_ms_per_epoch": train_ms, "epochs": E, "phases_ms": ph, "value": float(v), "torch_value": ref,
                  "rel_err": abs(float(v) - ref) / abs(ref), "grad_c": float(g_c),
                  }))
