"""BASELINE config 5 for real: Levy-100D, n=8192, 2^26-candidate pool sharded over the GPUs of one box,
GP fit on rank 0 only, one all-gather of per-GPU top-k.   torchrun --nproc-per-node N tools/run_config5_multigpu.py"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bbo_b200 as B
from bbo_b200.benchmarks import Levy, make_observations
from bbo_b200.distributed import broadcast_state, gather_topk, merge_topk_device, shard_range
from bbo_b200.pool import uniform_pool

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n, d, k = 8192, 100, 16
Q = int(os.environ.get("BBO_C5_POOL", str(1 << 26)))
X, y = make_observations(Levy(d), n, 11)
cov = B.ScaleKernel(B.RBFKernel(ard_dim=d)).double()
with torch.no_grad():
    cov.base_kernel.lengthscale.fill_(1.5); cov.variance.fill_(0.5413)
gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov, device=dev,
              noise_variance=1e-4, epochs=3, sign=-1.0, lr=0.01)
def sync():
    dist.barrier(); torch.cuda.synchronize()
sync(); t0 = time.perf_counter()
if rank == 0:
    gp.train()                                   # FP64 fit on ONE GPU (3 Adam epochs here)
    th = gp._theta.pack(dev)
else:
    th = torch.empty(gp._theta.P, dtype=torch.float64, device=dev)
torch.cuda.synchronize(); t_fit = time.perf_counter() - t0
dist.broadcast(th, src=0); gp._theta.unpack(th)
sync(); t0 = time.perf_counter()
broadcast_state(gp, src=0)                       # hyp, L^-1 (512 MB), alpha
gp._ensure_state(want_tf32=True)
sync(); t_bcast = time.perf_counter() - t0
acq = B.EI(float(y.max()))
lo, hi = shard_range(Q, rank, world)
res = {}
for prec in ("tf32",):
    sync(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run = None
    step = 1 << 22                                # generate + score the shard in 4 M-candidate slices (f32 pool)
    for s in range(lo, hi, step):
        q = min(step, hi - s)
        pool = uniform_pool(2026, s, q, d, device=dev, dtype=torch.float32)
        run = gp.score_pool(pool, acq, k, precision=prec, index_offset=s, running=run)
    gs, gi = gather_topk(*run)                   # the ONE collective
    top_s, top_i = merge_topk_device(gs, gi, k)
    e1.record(); sync()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    res[prec] = {"pool_seconds": float(ms) / 1e3, "cand_per_s": Q / (float(ms) / 1e3),
                 "top_index": int(top_i[0]), "top_score": float(top_s[0])}
    same = [torch.empty_like(top_i) for _ in range(world)]
    dist.all_gather(same, top_i)
    res[prec]["identical_on_all_ranks"] = all(torch.equal(t, top_i) for t in same)
if rank == 0:
    print(json.dumps({"config": 5, "desc": f"Levy-100D n=8192, pool {Q} over {world} GPU(s), on-device f32 pool, top-{k}",
                      "fit_3_epochs_s_rank0": t_fit, "broadcast_state_s": t_bcast, **res}), flush=True)
dist.destroy_process_group()
