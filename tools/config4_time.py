"""BASELINE config 4 on one B200: 256 independent GPs (n=256, d=10), one fit iteration (value + gradient) and one
Adam epoch inside train(), CUDA events.  BBO_FIT_SMALL=0 runs the multi-kernel pipeline for comparison.
Algorithmic work per iteration: B (n^3 + 3 n^2 d) FP64 flop."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bbo_b200.batch import CudaGPBatch

B, n, d = (int(v) for v in (sys.argv[1:4] or (256, 256, 10)))
rng = np.random.default_rng(0)
X = rng.random((B, n, d))
y = np.sin(3 * X.sum(-1)) + 0.1 * rng.standard_normal((B, n))
y = (y - y.mean(1, keepdims=True)) / y.std(1, keepdims=True)
theta = np.concatenate([np.zeros((B, 1)), rng.normal(-0.2, 0.1, (B, d)), np.full((B, 1), 0.5)], axis=1)
gb = CudaGPBatch(torch.as_tensor(X), torch.as_tensor(y)[..., None], "rbf_vmap", theta=torch.as_tensor(theta),
                 noise_variance=1e-4, epochs=50, lr=0.01)
for _ in range(3):
    gb.value_and_grad()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    v, g = gb.value_and_grad()
e1.record()
torch.cuda.synchronize()
vg_ms = e0.elapsed_time(e1) / 10
gb.train()
torch.cuda.synchronize()
gb.theta.copy_(torch.as_tensor(theta))
e0.record()
gb.train()
e1.record()
torch.cuda.synchronize()
ep_ms = e0.elapsed_time(e1) / 50
flop = B * (float(n) ** 3 + 3.0 * n * n * d)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                    "measured_peaks_tf32_f64.json")))
print(json.dumps({"config": 4, "B": B, "n": n, "d": d, "fit_small": os.environ.get("BBO_FIT_SMALL", "default"),
                  "value_grad_ms": vg_ms, "train_ms_per_epoch": ep_ms, "flop_per_iter": flop,
                  "tflops_value_grad": flop / vg_ms / 1e9, "tflops_epoch": flop / ep_ms / 1e9,
                  "frac_of_f64_peak_epoch": flop / ep_ms / 1e9 / peaks["f64_tflops"],
                  "value0": float(v[0]), "info_nonzero": int((gb.info != 0).sum())}))
