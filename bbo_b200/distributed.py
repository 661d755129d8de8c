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
    if mine.is_cuda:
        dist.all_gather_into_tensor(out, mine, group=group)   # one ncclAllGather straight into the packed buffer
    else:
        dist.all_gather(list(out.unbind(0)), mine, group=group)   # gloo (CPU tests of the host logic)
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


def broadcast_state(gp, src: int = 0, group=None, mode: str = "hyper") -> None:
    """Replicate the fitted model of `gp` from rank `src`.

    mode="hyper" (default): ONE broadcast of the raw hyper-parameter vector (1 + d + 1 doubles); every rank then
    factorises (X, y) once itself -- a single Cholesky + triangular inverse, 1/epochs of what the fit on `src` cost,
    identical bits on identical GPUs -- instead of receiving 8 np^2 bytes of L^-1 (537 MB at n = 8192: 40 ms
    measured in round 1 against ~17 ms for the local factorisation).  Every rank must hold the same X, y.
    mode="state": one fused broadcast of (hyp, alpha, L^-1) for callers whose ranks do not hold the data's
    factorisation inputs in float64-identical form."""
    rank = dist.get_rank(group)
    dev = gp._device
    if mode == "hyper":
        theta = gp._theta.pack(dev)
        dist.broadcast(theta, src=src, group=group)
        if rank != src:
            gp._theta.unpack(theta)
            gp._state = None
        gp.state        # factorise (src: no-op if already done)
        return
    if mode != "state":
        raise ValueError("mode must be 'hyper' or 'state'")
    nh, na, nl = 2 * gp.d + 2, gp.np, gp.np * gp.np
    buf = torch.empty(nh + na + nl, dtype=torch.float64, device=dev)
    if rank == src:
        s = gp.state
        buf[:nh].copy_(s["hyp"])
        buf[nh:nh + na].copy_(s["alpha"])
        buf[nh + na:].copy_(s["linv"].reshape(-1))
    dist.broadcast(buf, src=src, group=group)
    if rank != src:
        gp.install_state(buf[:nh], buf[nh + na:].view(gp.np, gp.np), buf[nh:nh + na])


def allgather_merge_topk(scores: torch.Tensor, indices: torch.Tensor, k: int, group=None):
    """The multi-GPU step on the device in two launches around ONE collective: bbo_topk_pack -> a single
    all_gather_into_tensor of the packed (score, index) pairs -> bbo_topk_merge_packed.  Returns the global
    (scores[k], indices[k]) on every rank, bit-identical for any number of GPUs."""
    lib = _lib.load()
    dev = _lib.require_cuda(scores.device)
    world = dist.get_world_size(group)
    kk = scores.numel()
    packed = torch.empty(2 * kk, dtype=torch.int64, device=dev)
    gathered = torch.empty(world * 2 * kk, dtype=torch.int64, device=dev)
    out_s = torch.empty(k, dtype=torch.float64, device=dev)
    out_i = torch.empty(k, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.stream_ptr(dev)
        _lib.check(lib.bbo_topk_pack(_lib.ptr(scores), _lib.ptr(indices), kk, _lib.ptr(packed), st), "bbo_topk_pack")
        dist.all_gather_into_tensor(gathered, packed, group=group)
        _lib.check(lib.bbo_topk_merge_packed(_lib.ptr(gathered), world * kk, k, _lib.ptr(out_s), _lib.ptr(out_i), st),
                   "bbo_topk_merge_packed")
    return out_s, out_i


def sharded_score_pool(gp, acq, k: int, Xq_local: torch.Tensor, index_offset: int, *, precision: str = "fp64",
                       group=None):
    """Score this rank's shard, all-gather the per-rank top-k, merge.  Returns the global
    (scores[k], indices[k]) on every rank."""
    s, i = gp.score_pool(Xq_local, acq, k, precision=precision, index_offset=index_offset)
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return s, i
    return allgather_merge_topk(s, i, k, group)
