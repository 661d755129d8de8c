"""Batched regularised-evolution multi-start acquisition optimiser.

The reference's default acquisition optimiser is ``DesignerAsOptimizer(RegularizedEvolutionDesigner1,
batch_size=1, num_evaluations=100)`` (bo/bo.py:26-33): one population of 25, tournament of 5, one
mutated coordinate per child (evolution/re1.py:31-114), 100 *sequential* single-candidate
``predict`` calls.  Here R independent populations (restarts) evolve in lock-step on the device:

  generation 0   pop_size random members per restart, scored in ONE fused call (the reference's
                 first pop_size suggestions are random too: re1.py:100-101)
  generation g   per restart: tournament_selection (re1.py:31-46: `tournament_size` distinct members,
                 best wins) -> random_mutation (re1.py:49-76: `num_mutation` distinct coordinates
                 resampled uniformly in [0,1)) -> all R children scored in one fused call ->
                 appended to the population, which is a FIFO of length pop_size
                 (collections.deque(maxlen=pop_size), re1.py:96,116-117)
  result         the `count` best of everything evaluated (optimizers/designer_optimizer.py:42,
                 get_best_trials order: descending score, ties -> earliest evaluation).

Population bookkeeping is a handful of tiny torch ops per generation (device agnostic, tested on
CPU); every score comes from the CUDA GP/acquisition kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .abstractions import GradientFreeOptimizer


def tournament_winner(fitness: torch.Tensor, cand: torch.Tensor) -> torch.Tensor:
    """re1.py:38-45 for R populations at once: fitness (R, P), drawn member indices cand (R, t) -> index of the
    winner (R,), the FIRST best of the drawn members in draw order (np.argmax semantics; goal MAXIMIZE: the
    acquisition problem is a maximisation, bo.py:90-92)."""
    f = fitness.gather(1, cand)
    best = f.max(dim=1, keepdim=True).values
    first = (f == best).to(torch.int8).argmax(dim=1, keepdim=True)     # first occurrence of the maximum
    return cand.gather(1, first).squeeze(1)


def tournament_select(fitness: torch.Tensor, tournament_size: int, gen: torch.Generator) -> torch.Tensor:
    """fitness (R, P) -> winner index (R,): best of `tournament_size` DISTINCT random members
    (rng.choice(..., replace=False), re1.py:37)."""
    R, P = fitness.shape
    keys = torch.rand(R, P, generator=gen, device=fitness.device)
    cand = keys.topk(min(tournament_size, P), dim=1).indices           # distinct by construction
    return tournament_winner(fitness, cand)


def apply_mutation(parents: torch.Tensor, cols: torch.Tensor, vals: torch.Tensor) -> torch.Tensor:
    """re1.py:57-64 for double parameters scaled to [0,1]: coordinates `cols` (R, k) of the parents are replaced by
    the uniform draws `vals` (R, k) (lb = 0, ub = 1: value = rng.random() * (ub - lb) + lb)."""
    return parents.scatter(1, cols, vals)


def mutate(parents: torch.Tensor, num_mutation: int, gen: torch.Generator) -> torch.Tensor:
    """parents (R, d) in [0,1): resample `num_mutation` distinct coordinates uniformly."""
    R, d = parents.shape
    keys = torch.rand(R, d, generator=gen, device=parents.device)
    cols = keys.topk(min(num_mutation, d), dim=1).indices
    vals = torch.rand(R, cols.shape[1], generator=gen, device=parents.device, dtype=parents.dtype)
    return apply_mutation(parents, cols, vals)


class RingPopulation:
    """R FIFO populations of capacity P (deque(maxlen=P) semantics)."""

    def __init__(self, X: torch.Tensor, F: torch.Tensor):
        self.X, self.F = X.clone(), F.clone()          # (R, P, d), (R, P)
        self.head = 0                                   # next slot to overwrite == oldest member

    def push(self, x: torch.Tensor, f: torch.Tensor) -> None:
        self.X[:, self.head] = x
        self.F[:, self.head] = f
        self.head = (self.head + 1) % self.X.shape[1]


class EvolutionMultiStart(GradientFreeOptimizer):
    def __init__(self, num_restarts: int = 64, pop_size: int = 25, tournament_size: int = 5,
                 num_mutation: int = 1, num_evaluations: int = 100, seed: int = 0):
        self.num_restarts, self.pop_size = int(num_restarts), int(pop_size)
        self.tournament_size, self.num_mutation = int(tournament_size), int(num_mutation)
        self.num_evaluations, self.seed = int(num_evaluations), int(seed)
        self._calls = 0

    def optimize_tensor(self, score_fn, d: int, *, count: int = 1,
                        seed_candidates: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, ...]:
        """score_fn: (q,d) float64 device tensor -> (q,) scores.  Returns (X_best (count,d), scores,
        flat evaluation ids)."""
        dev = score_fn.surrogate._device if hasattr(score_fn, "surrogate") else torch.device("cuda")
        self._calls += 1
        gen = torch.Generator(device=dev)
        gen.manual_seed(self.seed + 104729 * self._calls)
        R, P = self.num_restarts, self.pop_size
        X0 = torch.rand(R, P, d, generator=gen, device=dev, dtype=torch.float64)
        if seed_candidates is not None and len(seed_candidates):
            sc = seed_candidates.to(device=dev, dtype=torch.float64)[:P]
            X0[:, :sc.shape[0]] = sc
        F0 = score_fn(X0.reshape(R * P, d)).reshape(R, P)
        pop = RingPopulation(X0, F0)
        all_x, all_f = [X0.reshape(R * P, d)], [F0.reshape(-1)]
        for _ in range(max(self.num_evaluations - P, 0)):
            w = tournament_select(pop.F, self.tournament_size, gen)
            parents = pop.X[torch.arange(R, device=dev), w]
            children = mutate(parents, self.num_mutation, gen)
            f = score_fn(children)
            pop.push(children, f)
            all_x.append(children)
            all_f.append(f)
        X = torch.cat(all_x)
        F = torch.cat(all_f)
        idx = select_topk(F, count)
        return X[idx], F[idx], idx

    def optimize(self, score_fn, problem_statement, *, count: int = 1, seed_candidates=tuple()):
        d = problem_statement if isinstance(problem_statement, int) else len(problem_statement)
        return self.optimize_tensor(score_fn, d, count=count)[0]


def select_topk(scores: torch.Tensor, k: int) -> torch.Tensor:
    """Indices of the k best scores in get_best_trials order (descending, ties -> lowest index),
    through bbo_topk_merge (segments of 2048, merged in levels)."""
    from . import _lib
    lib = _lib.load()
    dev = _lib.require_cuda(scores.device)
    s = scores.detach().to(torch.float64).contiguous()
    i = torch.arange(s.numel(), dtype=torch.int64, device=dev)
    k = min(k, s.numel())
    with torch.cuda.device(dev):
        while True:
            m = s.numel()
            nseg = (m + 2047) // 2048
            out_s = torch.empty(nseg * k, dtype=torch.float64, device=dev)
            out_i = torch.empty(nseg * k, dtype=torch.int64, device=dev)
            for g in range(nseg):
                lo, hi = g * 2048, min(m, (g + 1) * 2048)
                _lib.check(lib.bbo_topk_merge(_lib.ptr(s[lo:hi]), _lib.ptr(i[lo:hi]), hi - lo, k,
                                              _lib.ptr(out_s[g * k:]), _lib.ptr(out_i[g * k:]), _lib.stream_ptr(dev)),
                           "bbo_topk_merge")
            s, i = out_s, out_i
            if nseg == 1:
                break
    return i[i >= 0]
