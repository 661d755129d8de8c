"""n-sweep of the scoring pass on one B200: candidates/s and the tensor-roofline fraction of the WHOLE step for
n in {256 ... 8192} at d=20 (ARD RBF, l=0.5, noise=1e-4, EI, top-8), in the headline mode (tf32+refine), the plain
tensor-core mode and float64.  One JSON line per n (profiles/n_sweep_r02.jsonl).  The roofline ceiling of a step is
peak / (n^2 + 3nd + 4n) candidates/s with peak = the sustained dense bf16 GEMM rate (the panel multiplies 16-bit
operands: MEASURED_PEAKS.json, else profiles/measured_peaks_tf32_f64.json) resp. the stored cuBLAS f64 rate."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bbo_b200 as B
from bbo_b200.benchmarks import Ackley, make_observations
from bbo_b200.pool import uniform_pool

D = 20
peaks = json.load(open(os.path.join(ROOT, "profiles", "measured_peaks_tf32_f64.json")))
try:
    peaks["bf16_tflops_sustained"] = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:  # noqa: BLE001
    pass


def ev(fn, reps=3, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


for n in [int(v) for v in (sys.argv[1:] or [256, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192])]:
    X, y = make_observations(Ackley(D), n, 20251019)
    cov = B.ScaleKernel(B.RBFKernel(ard_dim=D)).double()
    with torch.no_grad():
        cov.base_kernel.lengthscale.fill_(float(np.log(np.expm1(0.5))))
        cov.variance.fill_(float(np.log(np.expm1(1.0))))
    gp = B.CudaGP(torch.as_tensor(X), torch.as_tensor(y).reshape(-1, 1), B.ConstantMean().double(), cov,
                  noise_variance=1e-4)
    acq = B.EI(float(y.max()))
    q = (1 << 22) if n <= 1024 else (1 << 21 if n <= 2048 else 1 << 20)
    pool = uniform_pool(3, 0, q, D)      # float64, > L2
    gp.state
    flop = float(n) * n + 3.0 * n * D + 4.0 * n
    row = {"n": n, "d": D, "pool": q, "flop_per_candidate": flop}
    for prec in ("tf32+refine", "tf32", "fp64"):
        if prec == "fp64" and n > 2048:
            qs = 1 << 18
            ms = ev(lambda: gp.score_pool(pool[:qs], acq, 8, precision=prec), reps=1, warm=1)
            rate = qs / ms * 1e3
        else:
            ms = ev(lambda: gp.score_pool(pool, acq, 8, precision=prec))
            rate = q / ms * 1e3
        key = prec.replace("+", "_")
        row[key + "_cand_per_s"] = rate
        peak = peaks["f64_tflops"] if prec == "fp64" else peaks["bf16_tflops_sustained"]
        row[key + "_tflops"] = rate * flop / 1e12
        row[key + "_frac_of_" + ("f64_peak" if prec == "fp64" else "bf16_sustained")] = rate * flop / 1e12 / peak
    s64, i64 = gp.score_pool(pool[:1 << 18], acq, 8, precision="fp64")
    sr, ir = gp.score_pool(pool[:1 << 18], acq, 8, precision="tf32+refine")
    row["refine_selection_matches_fp64"] = bool(torch.equal(ir, i64))
    row["refine_certified"] = gp.refine_certificate()["certified"]
    print(json.dumps(row), flush=True)
    del gp, pool
    torch.cuda.empty_cache()
