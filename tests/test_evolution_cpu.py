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


def test_same_draws_give_the_references_winner_and_child():
    """Reference-driven check: the UNMODIFIED tournament_selection / random_mutation
    (evolution/re1.py:31-76) are run with a scripted rng -- the member indices, the mutated coordinates and the uniform
    values are prescribed -- and the tensor forms must pick the same winner and build the same child."""
    import numpy as np
    import pytest
    from _refenv import require_reference
    require_reference()
    from bbo.algorithms.converters.core import DefaultTrialConverter
    from bbo.algorithms.designers.evolution.re1 import random_mutation, tournament_selection
    from bbo.shared.base_study_config import MetricInformation, ObjectiveMetricGoal, ProblemStatement
    from bbo.shared.trial import Measurement, Trial
    from bbo_b200.evolution import apply_mutation, tournament_winner

    d, P, t = 6, 25, 5
    ps = ProblemStatement()
    for k in range(d):
        ps.search_space.add_float_param(f"x{k}", -2.0, 3.0)
    ps.metric_information.append(MetricInformation("acquisition", ObjectiveMetricGoal.MAXIMIZE))
    conv = DefaultTrialConverter.from_study_config(ps)
    names = list(conv.output_specs.keys())
    rng = np.random.default_rng(0)
    U = rng.random((P, d))                                   # population in [0,1]^d (the converter's scaled features)
    fit = np.round(rng.standard_normal(P), 1)                # rounded: ties inside tournaments do occur
    pop = conv.to_trials({n: U[:, [k]].astype(np.float32) for k, n in enumerate(names)})
    for tr, f in zip(pop, fit):
        tr.complete(Measurement({"acquisition": float(f)}))
    U32 = np.concatenate([conv.to_features(pop)[n] for n in names], axis=1).astype(np.float64)

    class Scripted:
        """np.random.Generator stand-in returning prescribed draws."""
        def __init__(self, choices, values):
            self.choices, self.values = list(choices), list(values)
        def choice(self, a, size=None, replace=True):
            return self.choices.pop(0)
        def random(self, shape, dtype=np.float64):
            return np.full(shape, self.values.pop(0), dtype=dtype)

    for trial in range(40):
        cand = rng.choice(P, size=t, replace=False)
        winner = tournament_selection(pop, t, ObjectiveMetricGoal.MAXIMIZE, Scripted([cand], []))
        w = tournament_winner(torch.as_tensor(fit)[None, :], torch.as_tensor(cand)[None, :])
        assert pop[int(w[0])] is winner
        nm = 1 + trial % 2
        cols = rng.choice(d, size=nm, replace=False)
        vals = rng.random(nm).astype(np.float32)
        child = random_mutation(winner, conv, nm, Scripted([np.array(names)[cols]], vals))
        got = apply_mutation(torch.as_tensor(U32[[int(w[0])]]), torch.as_tensor(cols)[None, :],
                             torch.as_tensor(vals.astype(np.float64))[None, :])
        want = np.concatenate([conv.to_features([child])[n] for n in names], axis=1).astype(np.float64)
        np.testing.assert_allclose(got.numpy(), want, rtol=0, atol=2e-7)      # float32 feature round trip
