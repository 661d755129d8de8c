"""Multi-GPU pool scoring: one process per GPU, candidates sharded by contiguous index range.

The fitted state (hyper-parameters, L^-1, alpha) is replicated -- computed on rank 0 and broadcast
once per suggestion (north_star: "the GP fit itself runs on one GPU") -- and every rank scores its
own slice of the pool.  The only data-path collective is ONE all-gather of each rank's k
(score, index) pairs (k*16 bytes per rank: latency bound, bandwidth irrelevant), followed by the
same deterministic merge on every rank (descending score, ties -> lowest global index,
shared/trial.py:289-294), so the selection is bit-identical for any number of GPUs.
The reference has no distributed counterpart.
"""
from __future__ import annotations

import ctypes as C
from typing import Tuple

import torch
import torch.distributed as dist

from . import _lib


def shard_range(q_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the global pool owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(q_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_topk(scores: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """(k,) float64 + (k,) int64 -> (k,2) int64 (scores bit-cast) so ONE collective moves both."""
    return torch.stack([scores.contiguous().view(torch.int64), indices.contiguous()], dim=1).contiguous()


def unpack_topk(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    p = packed.reshape(-1, 2)
    return p[:, 0].contiguous().view(torch.float64), p[:, 1].contiguous()


def gather_topk(scores: torch.Tensor, indices: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The single data-path collective: all-gather every rank's k pairs -> (world*k,) each.
    Backend agnostic (NCCL on GPUs; gloo in the CPU tests of the host logic)."""
    world = dist.get_world_size(group)
    mine = pack_topk(scores, indices)
    out = torch.empty((world,) + tuple(mine.shape), dtype=torch.int64, device=mine.device)
    dist.all_gather(list(out.unbind(0)), mine, group=group)   # one ncclAllGather (gloo in CPU tests)
    return unpack_topk(out)


def merge_topk_device(scores: torch.Tensor, indices: torch.Tensor, k: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """k best of the gathered pairs on the device (bbo_topk_merge)."""
    lib = _lib.load()
    dev = _lib.require_cuda(scores.device)
    out_s = torch.empty(k, dtype=torch.float64, device=dev)
    out_i = torch.empty(k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.bbo_topk_merge(_lib.ptr(scores), _lib.ptr(indices), scores.numel(), k, _lib.ptr(out_s),
                                      _lib.ptr(out_i), _lib.stream_ptr(dev)), "bbo_topk_merge")
    return out_s, out_i


def broadcast_state(gp, src: int = 0, group=None) -> None:
    """Replicate the fitted state of `gp` from rank `src` (one broadcast per tensor; n=8192 is
    ~0.5 GB of float64 L^-1, a few ms over NVLink).  Ranks other than src skip the factorisation."""
    rank = dist.get_rank(group)
    if rank == src:
        s = gp.state
        hyp, linv, alpha = s["hyp"], s["linv"], s["alpha"]
    else:
        dev = gp._device
        hyp = torch.empty(2 * gp.d + 2, dtype=torch.float64, device=dev)
        linv = torch.empty(gp.np, gp.np, dtype=torch.float64, device=dev)
        alpha = torch.empty(gp.np, dtype=torch.float64, device=dev)
    for t in (hyp, linv, alpha):
        dist.broadcast(t, src=src, group=group)
    if rank != src:
        gp.install_state(hyp, linv, alpha)


def sharded_score_pool(gp, acq, k: int, Xq_local: torch.Tensor, index_offset: int, *, precision: str = "fp64",
                       group=None):
    """Score this rank's shard, all-gather the per-rank top-k, merge.  Returns the global
    (scores[k], indices[k]) on every rank."""
    s, i = gp.score_pool(Xq_local, acq, k, precision=precision, index_offset=index_offset)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return s, i
    gs, gi = gather_topk(s, i, group)
    return merge_topk_device(gs, gi, k)
