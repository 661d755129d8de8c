"""Ad-hoc timing of the main entry points (development aid; bench.py is the contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bbo_b200 as B

def ev_time(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)

def peaks():
    out = {}
    for name, dt, tf32 in [("fp64", torch.float64, False), ("tf32", torch.float32, True), ("fp32", torch.float32, False)]:
        torch.backends.cuda.matmul.allow_tf32 = tf32
        N = 8192 if name != "fp64" else 4096
        a = torch.randn(N, N, dtype=dt, device="cuda"); b = torch.randn(N, N, dtype=dt, device="cuda")
        ms = ev_time(lambda: a @ b, reps=5, warm=2)
        out[name + "_tflops"] = 2 * N**3 / ms / 1e9
    torch.backends.cuda.matmul.allow_tf32 = False
    return out

def main():
    res = {"peaks": peaks()}
    print(json.dumps(res), flush=True)
    for n, d in [(1024, 20), (4096, 20)]:
        rng = np.random.default_rng(0)
        X = torch.as_tensor(rng.random((n, d))); y = torch.as_tensor(np.sin(X.sum(1).numpy()*2)).reshape(-1, 1)
        y = (y - y.mean()) / y.std()
        cov = B.ScaleKernel(B.RBFKernel(ard_dim=d)).double()
        with torch.no_grad():
            cov.base_kernel.lengthscale.fill_(0.0); cov.variance.fill_(0.5)
        gp = B.CudaGP(X, y, B.ConstantMean().double(), cov, noise_variance=1e-4, epochs=5)
        t_vg = ev_time(lambda: gp.value_and_grad(), reps=3)
        t0 = time.time(); gp.train(); torch.cuda.synchronize(); t_train = (time.time() - t0) / 5
        q = 1 << 16
        Xq = torch.rand(q, d, dtype=torch.float64, device="cuda")
        acq = B.EI(float(y.max()))
        gp.state
        t_sc = ev_time(lambda: gp.score_pool(Xq, acq, 8), reps=3)
        flops = q * (n * n + 3 * n * d + 4 * n)
        r = {"n": n, "d": d, "value_grad_ms": t_vg, "train_ms_per_epoch": t_train * 1e3,
             "fit_tflops": (n**3 + 3 * n * n * d) / t_vg / 1e9,
             "score_fp64_ms": t_sc, "score_fp64_cand_per_s": q / t_sc * 1e3, "score_fp64_tflops": flops / t_sc / 1e9}
        print(json.dumps(r), flush=True)

def batched():
    from bbo_b200.batch import CudaGPBatch
    B_, n, d = 256, 256, 10
    rng = np.random.default_rng(1)
    X = torch.as_tensor(rng.random((B_, n, d))); y = torch.sin(X.sum(-1, keepdim=True) * 3)
    y = (y - y.mean(1, keepdim=True)) / y.std(1, keepdim=True)
    th = torch.zeros(B_, 1 + d + 1, dtype=torch.float64); th[:, 1:] = 0.3
    gb = CudaGPBatch(X, y, "matern52_vmap", theta=th, noise_variance=1e-4, epochs=10)
    t = ev_time(lambda: gb.value_and_grad(), reps=3)
    t0 = time.time(); gb.train(); torch.cuda.synchronize(); tt = (time.time() - t0) / 10
    Xq = torch.rand(B_, 4096, d, dtype=torch.float64, device="cuda")
    gb.score_pools(Xq, lambda b: B.EI(1.0), 4); torch.cuda.synchronize()
    t0 = time.time(); gb.score_pools(Xq, lambda b: B.EI(1.0), 4); torch.cuda.synchronize(); ts = time.time() - t0
    print(json.dumps({"config4": "256 x (n=256, d=10)", "value_grad_ms_all": t, "train_ms_per_epoch_all": tt * 1e3,
                      "fits_per_s": B_ / (t * 1e-3), "score_256x4096_s": ts, "cand_per_s": B_ * 4096 / ts}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "batched":
        batched(); sys.exit(0)
    main()
