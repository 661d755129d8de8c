"""Loader for tests/golden/*.npz (generated from the real reference by make_golden.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import gp_oracle as O  # noqa: E402

_gp = np.load(os.path.join(HERE, "golden", "gp_cases.npz"))
CASE_NAMES = [str(s) for s in _gp["names"]]


class Case:
    def __init__(self, name):
        g = lambda k: _gp[f"{name}/{k}"]
        self.name = name
        n, d, q, ard, scale, kind = [int(v) for v in g("meta")]
        self.n, self.d, self.q, self.ard, self.scale = n, d, q, bool(ard), bool(scale)
        self.kind = O.RBF if kind == 0 else O.MATERN52
        self.noise, self.eps = [float(v) for v in g("noise_eps")]
        for k in ["X", "Y", "Xq", "raw_ls", "raw_var", "const", "Kxx", "Kqx", "value", "g_ls", "g_var",
                  "g_c", "mu", "var_diag", "var_full", "chol_diag", "kinvy", "train_raw_ls",
                  "train_raw_var", "train_const", "train_mu", "train_var_diag"]:
            setattr(self, k, g(k))

    def params(self, trained=False):
        return O.GPParams(
            kind=self.kind,
            raw_lengthscale=np.array(self.train_raw_ls if trained else self.raw_ls, dtype=np.float64),
            raw_variance=(float(self.train_raw_var if trained else self.raw_var) if self.scale else None),
            constant=float(self.train_const if trained else self.const),
            noise=self.noise, pairwise_eps=self.eps)


def load_acq():
    return np.load(os.path.join(HERE, "golden", "acq_cases.npz"))


def load_select():
    return np.load(os.path.join(HERE, "golden", "select_cases.npz"))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))) if a.size else 0.0
