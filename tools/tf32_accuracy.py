"""TF32-mode accuracy against the FP64 mode of the same library (development aid): max / mean signed error of mu and
var for a few shapes.  Run twice (default, BBO_KSTAR_NO_UMMA=1) to compare K* producers."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bbo_b200 as B
from _golden import O
from _gpu_helpers import gp_from_params, synth_problem

for n, d, kind, ls in [(254, 12, O.RBF, -0.2), (1024, 20, O.RBF, 0.2), (2100, 20, O.RBF, 0.2), (600, 24, O.MATERN52, 0.2),
                       (200, 4, O.RBF, -2.0), (4096, 20, O.RBF, 0.2)]:
    p, X, y = synth_problem(1000 + n, n, d, kind=kind, noise=1e-4, ls_mean=ls)
    Xq = torch.as_tensor(np.random.default_rng(n).random((4096, d))).cuda()
    gp = gp_from_params(p, X, y)
    acq = B.EI(float(y.max()))
    _, _, mu64, var64, s64 = gp.score_pool(Xq, acq, 4, precision="fp64", return_all=True)
    _, _, mu32, var32, s32 = gp.score_pool(Xq, acq, 4, precision="tf32", return_all=True)
    em, ev, es = (mu32 - mu64).cpu().numpy(), (var32 - var64).cpu().numpy(), (s32 - s64).cpu().numpy()
    print(json.dumps({"n": n, "d": d, "kind": kind, "mu_max": float(np.abs(em).max()), "mu_mean": float(em.mean()),
                      "var_max": float(np.abs(ev).max()), "var_mean": float(ev.mean()),
                      "score_max": float(np.abs(es).max()), "score_scale": float(s64.abs().max())}), flush=True)
