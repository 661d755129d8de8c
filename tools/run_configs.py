"""Runs the five BASELINE.json configurations at their real sizes on one GPU (config 5: the per-GPU
shard of an 8-way split) and prints one JSON line each.  Development / evidence script."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B
from bbo_b200.benchmarks import Ackley, Branin2D, Levy, Rastrigin, make_observations
from bbo_b200.pool import uniform_pool

def ev(fn, reps=2, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); best = min(best, a.elapsed_time(b))
    return best

def gp_for(func, n, kind, noise=1e-4, raw_ls=0.0, epochs=100, **kw):
    X, y = make_observations(func, n, 7)
    base = (B.RBFKernel if kind == "rbf" else B.Matern52Kernel)(ard_dim=func.dim)
    cov = B.ScaleKernel(base).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(raw_ls); cov.variance.fill_(0.5413)
    gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
                  noise_variance=noise, epochs=epochs, **kw)
    return gp, X, y

out = []
# C1: GP-BO (RBF, EI) on 2-D Branin, 100 observations, 8192 random candidates
gp, X, y = gp_for(Branin2D(), 100, "rbf", noise=1e-4, epochs=100, sign=-1.0, lr=0.05)
t_train = ev(lambda: gp.train(), reps=1, warm=1)
pool = uniform_pool(1, 0, 8192, 2)
acq = B.EI(float(y.max()))
t_sc = ev(lambda: gp.score_pool(pool, acq, 1))
out.append({"config": 1, "desc": "Branin-2D n=100 q=8192 RBF EI", "train_100_epochs_ms": t_train, "score_ms": t_sc,
            "suggest_ms": t_train + t_sc, "cand_per_s": 8192 / t_sc * 1e3})
# C2: Ackley-20D, n=1024, 2^20 pool
gp, X, y = gp_for(Ackley(20), 1024, "rbf", raw_ls=0.0)
pool = uniform_pool(2, 0, 1 << 20, 20); acq = B.EI(float(y.max())); gp.state
r = {"config": 2, "desc": "Ackley-20D n=1024 q=2^20 EI"}
for prec in ("fp64", "tf32", "tf32+refine"):
    t = ev(lambda: gp.score_pool(pool, acq, 8, precision=prec)); r[prec + "_ms"] = t; r[prec + "_cand_per_s"] = (1 << 20) / t * 1e3
r["refine_matches_fp64"] = bool(torch.equal(gp.score_pool(pool, acq, 8, precision="tf32+refine")[1], gp.score_pool(pool, acq, 8)[1]))
out.append(r)
# C3: Rastrigin-50D, Matern-5/2 ARD, n=4096, FP64 fit + 64 multi-start restarts (as 64 pool shards)
gp, X, y = gp_for(Rastrigin(50), 4096, "matern52", raw_ls=1.0, noise=1e-4)
t_fit = ev(lambda: gp.value_and_grad(), reps=2)
v, g_ls, g_var, g_c = gp.value_and_grad()
acq = B.EI(float(y.max())); gp.state
pools = torch.cat([uniform_pool(100 + s, 0, 4096, 50) for s in range(64)])      # 64 restarts x 4096, one fused call
r = {"config": 3, "desc": "Rastrigin-50D Matern52 ARD n=4096: fit + 64 restarts x 4096 (one fused call) + "
                          "64-restart regularised evolution (25 pop, 100 evaluations each)",
     "fit_ms_per_iter": t_fit, "value": float(v), "grad_finite": bool(torch.isfinite(g_ls).all())}
for prec in ("fp64", "tf32", "tf32+refine"):
    t = ev(lambda: gp.score_pool(pools, acq, 4, precision=prec), reps=1); r[prec + "_64x4096_ms"] = t
    r[prec + "_cand_per_s"] = 64 * 4096 / t * 1e3
from bbo_b200.evolution import EvolutionMultiStart
from bbo_b200.optimizers import TensorScoreFn
class _Fn(TensorScoreFn):
    def __call__(self, X):
        return self.surrogate.score_pool(X, self.acquisition, 1, precision=self.precision, return_all=True)[4]
for prec in ("fp64", "tf32"):
    opt = EvolutionMultiStart(num_restarts=64, pop_size=25, num_evaluations=100, seed=1)
    fn = _Fn(gp, acq, prec)
    t = ev(lambda: opt.optimize_tensor(fn, 50, count=4), reps=1, warm=1)
    r["evolution_64x100_" + prec + "_ms"] = t
out.append(r)
# C4: 256 independent fits (n=256, d=10) scored concurrently
from bbo_b200.batch import CudaGPBatch
rng = np.random.default_rng(1)
Xb = torch.as_tensor(rng.random((256, 256, 10))); yb = torch.sin(Xb.sum(-1, keepdim=True) * 3)
yb = (yb - yb.mean(1, keepdim=True)) / yb.std(1, keepdim=True)
th = torch.zeros(256, 12, dtype=torch.float64); th[:, 1:] = 0.3
gb = CudaGPBatch(Xb, yb, "matern52_vmap", theta=th, noise_variance=1e-4, epochs=50)
t = ev(lambda: gb.value_and_grad())
gb.train(); torch.cuda.synchronize(); gb.theta.copy_(th)
t_ep = ev(lambda: gb.train(), reps=1, warm=0) / 50
Xq = torch.rand(256, 4096, 10, dtype=torch.float64, device="cuda")
gb.score_pools(Xq, lambda b: B.EI(1.0), 4); torch.cuda.synchronize()
t0 = time.time(); gb.score_pools(Xq, lambda b: B.EI(1.0), 4); torch.cuda.synchronize(); ts = time.time() - t0
out.append({"config": 4, "desc": "256 x (n=256,d=10) fits + 256 x 4096 scoring", "value_grad_ms_all_256": t,
            "us_per_fit_iter": t * 1e3 / 256, "train_ms_per_epoch_all_256": t_ep, "score_all_s": ts, "cand_per_s": 256 * 4096 / ts})
# C5: Levy-100D, n=8192, 2^26 pool over 8 GPUs -> one GPU's shard is 2^23; time 2^20 of it
gp, X, y = gp_for(Levy(100), 8192, "rbf", raw_ls=1.5, noise=1e-4)
t_fit = ev(lambda: gp.value_and_grad(), reps=1)
acq = B.EI(float(y.max())); gp.state
pool = uniform_pool(5, 0, 1 << 20, 100, dtype=torch.float32)
r = {"config": 5, "desc": "Levy-100D n=8192: fit iter + 2^20-candidate slice of the 2^26 pool (f32 candidates)",
     "fit_ms_per_iter": t_fit}
for prec, q in (("fp64", 1 << 17), ("tf32", 1 << 20), ("tf32+refine", 1 << 20)):
    t = ev(lambda: gp.score_pool(pool[:q], acq, 8, precision=prec), reps=1); r[prec + "_ms"] = t
    r[prec + "_cand_per_s"] = q / t * 1e3
r["refine_full_pool_8gpu_s_est"] = (1 << 26) / 8 / r["tf32+refine_cand_per_s"]
out.append(r)
# N3: appending 8 observations to a factorised n=4096 state vs refactorising
gp, X, y = gp_for(Ackley(20), 4096 + 64, "rbf", raw_ls=0.0)
Xt, yt = torch.as_tensor(X).cuda(), torch.as_tensor(y).reshape(-1, 1).cuda()
def fresh():
    g = B.CudaGP(Xt[:4096], yt[:4096], B.ConstantMean().double(), gp._cov_module, noise_variance=1e-4)
    g.state
    return g
g0 = fresh(); torch.cuda.synchronize()
t_fact = ev(lambda: fresh(), reps=2, warm=0)
g0.append(Xt[4096:4104], yt[4096:4104]); torch.cuda.synchronize()          # warm-up (re-layout 4096 -> 4224 rows)
ts = []
for i in range(3):                                                          # wall clock incl. the info read-back
    t0 = time.perf_counter()
    g0.append(Xt[4104 + 8 * i:4112 + 8 * i], yt[4104 + 8 * i:4112 + 8 * i]); torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
out.append({"row": "N3", "desc": "append 8 observations at n~4104 (Ackley-20D) vs factorise from scratch",
            "append_ms_wall": min(ts), "factorize_ms": t_fact, "fit_100_epochs_ms_est": 100 * 4.5})
# N4: one objective value + full gradient through a Kumaraswamy-warped Matern kernel at n=4096, d=20
cov = B.ScaleKernel(B.WarpKernel(B.Matern52Kernel(ard_dim=20), B.KumarWarp())).double()
wg = B.CudaWarpedGP(Xt[:4096], yt[:4096], B.ConstantMean().double(), cov, noise_variance=1e-4)
def vg():
    for p_ in cov.parameters(): p_.grad = None
    wg.objective().backward()
t_w = ev(vg, reps=2)
pool = uniform_pool(9, 0, 1 << 18, 20, dtype=torch.float32)
acq = B.EI(float(y.max()))
t_ws = ev(lambda: wg.score_pool(pool, acq, 8, precision="tf32"), reps=2)
out.append({"row": "N4", "desc": "KumarWarp + Matern52 ARD, n=4096, d=20", "value_grad_all_params_ms": t_w,
            "tf32_score_2^18_ms": t_ws, "tf32_cand_per_s": (1 << 18) / t_ws * 1e3})
for o in out:
    print(json.dumps(o), flush=True)
