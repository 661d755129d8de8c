"""Make the unmodified reference importable as ``bbo`` (tests, bench.py reference arm, drop-in tests).

Search order: $BBO_REFERENCE, <repo>/baseline/_ref (vendored copy, travels to the GPU box), /root/reference
(build container).  ``bbo.algorithms.designers`` imports ``GPyTorchGP`` (designers/__init__.py:4 ->
bo/bo.py:12), and gpytorch is not in the image; when it is missing, a stub of exactly the
names imported at gpytorch_gp.py:6-13 / gpytorch_gp/kernels.py:2 is pre-seeded in sys.modules.  The stub
carries no arithmetic -- the default GPyTorchGP surrogate stays out of scope (parity unpinned).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = (os.environ.get("BBO_REFERENCE"), os.path.join(HERE, "_ref"), "/root/reference")


def reference_path():
    for p in CANDIDATES:
        if p and os.path.isdir(os.path.join(p, "bbo")):
            return p
    return None


def install_gpytorch_stub() -> bool:
    """Returns True if the stub was installed (i.e. the real gpytorch is absent)."""
    if "gpytorch" in sys.modules:
        return getattr(sys.modules["gpytorch"], "__bbo_stub__", False)
    try:
        importlib.import_module("gpytorch")
        return False
    except Exception:  # noqa: BLE001
        pass
    import torch

    wanted = {
        "gpytorch": [],
        "gpytorch.models": ["ExactGP"],
        "gpytorch.constraints": [],
        "gpytorch.constraints.constraints": ["Interval"],
        "gpytorch.distributions": ["MultivariateNormal"],
        "gpytorch.means": ["ConstantMean"],
        "gpytorch.kernels": ["RBFKernel", "MaternKernel", "ScaleKernel", "Kernel"],
        "gpytorch.likelihoods": ["GaussianLikelihood"],
        "gpytorch.mlls": ["ExactMarginalLogLikelihood"],
        "gpytorch.priors": ["torch_priors"],
    }
    for name, attrs in wanted.items():
        m = types.ModuleType(name)
        m.__path__ = []
        m.__bbo_stub__ = True
        for a in attrs:
            setattr(m, a, type(a, (torch.nn.Module,), {}))
        if name == "gpytorch.priors":    # module-level calls at gpytorch_gp.py:22-26
            m.torch_priors = types.SimpleNamespace(GammaPrior=lambda *a, **k: None,
                                                   LogNormalPrior=lambda *a, **k: None)
        sys.modules[name] = m
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, m)
    return True


def setup_reference(required: bool = True):
    """Put the reference on sys.path (+ gpytorch stub if needed).  Returns its path or None."""
    p = reference_path()
    if p is None:
        if required:
            raise RuntimeError("reference package `bbo` not found: run `python baseline/vendor_reference.py` "
                               "in the build container (baseline/_ref travels to the GPU box)")
        return None
    if p not in sys.path:
        sys.path.insert(0, p)
    install_gpytorch_stub()
    return p
