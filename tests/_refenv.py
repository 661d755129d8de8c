"""Reference (`bbo`) access for the drop-in tests: vendored copy under baseline/_ref (travels to the GPU box) or
/root/reference in the build container.  Plus a test-side adapter that gives the UNMODIFIED reference GP the
(q,1,d) -> (q,1,1) predict convention BODesigner.score_fn needs (its own 3-D path is broken)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from baseline.ref_env import setup_reference  # noqa: E402


def require_reference():
    p = setup_reference(required=False)
    if p is None:
        pytest.skip("reference package not available (run baseline/vendor_reference.py)")
    return p


class RefGPAdapter:
    """Wraps the reference's native GP (surrogates/gp/gp.py:30-83, float64 on the CPU): 2-D chunks -> diagonal."""

    def __init__(self, gp, chunk=512):
        self.gp = gp
        self.chunk = chunk

    def train(self):
        self.gp.train()

    def predict(self, X):
        if X.dim() == 2:
            return self.gp.predict(X.double().cpu())
        q = X.shape[0]
        x2 = X.reshape(q, -1).double().cpu()
        mus, vars_ = [], []
        with torch.no_grad():
            for a in range(0, q, self.chunk):
                mu, cov = self.gp.predict(x2[a:a + self.chunk])
                mus.append(mu.reshape(-1))
                vars_.append(cov.diagonal())
        return torch.cat(mus).reshape(q, 1, 1), torch.cat(vars_).reshape(q, 1, 1)


def register_adapter():
    from bbo.algorithms.abstractions import Surrogate
    Surrogate.register(RefGPAdapter)
