"""torchrun --nproc-per-node N tools/multigpu_check.py : the N-GPU sharded selection must be
bit-identical to the 1-GPU selection on the same global pool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bbo_b200 as B
from bbo_b200.distributed import broadcast_state, sharded_score_pool, shard_range
from bbo_b200.pool import uniform_pool
from bbo_b200.benchmarks import Levy, make_observations

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, q_total, k = 1024, 30, 200_003, 16
X, y = make_observations(Levy(d), n, 3)
cov = B.ScaleKernel(B.Matern52Kernel(ard_dim=d)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(0.8); cov.variance.fill_(0.5)
gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov, device=dev,
              noise_variance=1e-4, epochs=5, sign=-1.0)
if rank == 0:
    gp.train()                      # fit on one GPU only
    th = gp._theta.pack(dev)
else:
    th = torch.empty(gp._theta.P, dtype=torch.float64, device=dev)
dist.broadcast(th, src=0)           # raw hyper-parameters (tiny) so every rank builds the same hyp
gp._theta.unpack(th)
broadcast_state(gp, src=0)
acq = B.EI(float(y.max()))
ok = True
for prec in ("fp64", "tf32"):
    lo, hi = shard_range(q_total, rank, world)
    pool = uniform_pool(99, lo, hi - lo, d, device=dev)
    s, i = sharded_score_pool(gp, acq, k, pool, lo, precision=prec)
    if rank == 0:
        full = uniform_pool(99, 0, q_total, d, device=dev)
        s1, i1 = gp.score_pool(full, acq, k, precision=prec)
        same = torch.equal(i, i1) and torch.equal(s, s1)
        print(f"[{prec}] world={world} sharded == single-GPU selection: {same}; top index {int(i[0])} score {float(s[0]):.6e}")
        ok = ok and same
    gathered = [torch.empty_like(i) for _ in range(world)]
    dist.all_gather(gathered, i)
    if rank == 0:
        ok = ok and all(torch.equal(g, i) for g in gathered)
if rank == 0:
    print("MULTIGPU_CHECK", "PASS" if ok else "FAIL")
dist.destroy_process_group()
