"""Candidate pools on the device.

``uniform_pool`` replaces RandomDesigner.suggest (bbo/algorithms/designers/random.py:41-57) for the
acquisition pool: a counter-based Philox generator whose values depend only on
(seed, global row, column), so that every sharding of the pool (chunks, GPUs) sees the same
candidates and no Python ``Trial`` object is ever created for a pool member.
"""
from __future__ import annotations

import torch

from . import _lib


def uniform_pool(seed: int, row_offset: int, q: int, d: int, *, device="cuda", dtype=torch.float64,
                 out: torch.Tensor | None = None) -> torch.Tensor:
    lib = _lib.load()
    dev = _lib.require_cuda(device)
    if dtype not in (torch.float64, torch.float32):
        raise ValueError("dtype must be float64 or float32")
    if out is None:
        out = torch.empty(q, d, dtype=dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.bbo_pool_uniform(int(seed) & (2**64 - 1), int(row_offset), q, d, _lib.ptr(out),
                                        int(dtype == torch.float32), _lib.stream_ptr(dev)), "bbo_pool_uniform")
    return out
