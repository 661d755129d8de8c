"""Loader for tests/golden/warp_cases.npz (reference outputs for WarpKernel / KumarWarp / MLPWarp / MLPMean)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_w = np.load(os.path.join(HERE, "golden", "warp_cases.npz"))
WARP_CASES = ["kumar", "mlp"]


class WarpCase:
    def __init__(self, name):
        self.name = name
        self.n, self.d, self.q = [int(v) for v in _w[name + "/meta"]]
        self.noise = float(_w[name + "/noise"])
        for k in ["X", "Y", "Xq", "value", "mu", "var_diag", "train_mu", "train_var_diag"]:
            setattr(self, k, _w[f"{name}/{k}"])

    def group(self, prefix):
        """{state-dict key: array} for prefix in mean, cov, gmean, gcov, tmean, tcov."""
        p = f"{self.name}/{prefix}/"
        return {k[len(p):]: _w[k] for k in _w.files if k.startswith(p)}
