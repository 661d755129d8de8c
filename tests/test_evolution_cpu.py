"""CPU: population bookkeeping of the batched regularised-evolution optimiser
(semantics of evolution/re1.py:31-76,96,116-117 of the reference)."""
import torch

from bbo_b200.evolution import RingPopulation, mutate, tournament_select


def test_tournament_picks_best_of_distinct_members():
    gen = torch.Generator().manual_seed(0)
    F = torch.arange(25, dtype=torch.float64).repeat(64, 1)          # fitness == member index
    for _ in range(20):
        w = tournament_select(F, 5, gen)
        assert w.shape == (64,) and (w >= 4).all()                   # best of 5 distinct members is >= 4
    # tournament of the whole population always returns the arg-max
    F2 = torch.randn(8, 25, generator=gen, dtype=torch.float64)
    assert torch.equal(tournament_select(F2, 25, gen), F2.argmax(1))
    # the winner distribution favours fit members: mean winner index well above the uniform mean (12)
    ws = torch.cat([tournament_select(F, 5, gen) for _ in range(50)]).double()
    assert ws.mean() > 18


def test_mutation_changes_exactly_k_coordinates_in_bounds():
    gen = torch.Generator().manual_seed(1)
    P = torch.rand(64, 20, generator=gen, dtype=torch.float64)
    for k in (1, 3):
        C = mutate(P, k, gen)
        changed = (C != P).sum(1)
        assert (changed == k).all()
        assert (C >= 0).all() and (C < 1).all()


def test_ring_population_is_fifo():
    X = torch.zeros(2, 3, 1, dtype=torch.float64)
    F = torch.tensor([[0., 1., 2.], [0., 1., 2.]], dtype=torch.float64)
    pop = RingPopulation(X, F)
    for t in range(4):
        pop.push(torch.full((2, 1), 10.0 + t, dtype=torch.float64), torch.full((2,), 10.0 + t, dtype=torch.float64))
    # capacity 3, 4 pushes: the three youngest (11, 12, 13) survive, like deque(maxlen=3)
    assert sorted(pop.F[0].tolist()) == [11.0, 12.0, 13.0]
