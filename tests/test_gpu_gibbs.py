"""Gibbs sampling on the device (BASELINE.json configs[4]; reference: bayes_net.py:665-737).

The reference's chain cannot be reproduced draw for draw (its RNG is Python's `random` plus a
Cython alias sampler), so parity is distributional: the chain's frequencies must converge to
the exact posterior -- which is itself pinned to the reference -- and the run must be
reproducible from the seed."""
import numpy as np
import pandas as pd
import pytest

pytestmark = pytest.mark.gpu


def test_gibbs_converges_to_exact_posterior_on_reference_examples():
    from sorobn_b200 import examples

    cases = [
        (examples.sprinkler, ("Rain",), {"Sprinkler": True}),            # bayes_net.py:683-688
        (examples.alarm, ("Burglary",), {"John calls": True, "Mary calls": True}),
        (examples.grades, ("Letter", "SAT"), {"Intelligence": "Smart"}),
        (examples.sprinkler, ("Cloudy", "Rain"), {}),
    ]
    for make, query, event in cases:
        bn = make(seed=42)
        exact = bn.query(*query, event=event)
        got = bn.query(*query, event=event, algorithm="gibbs", n_iterations=400_000)
        assert got.name == exact.name and list(got.index.names) == list(exact.index.names)
        assert np.isclose(got.sum(), 1.0, atol=1e-5)
        both = pd.concat([exact, got], axis=1).fillna(0.0)
        assert np.abs(both.iloc[:, 0] - both.iloc[:, 1]).max() < 0.01, both


def test_gibbs_is_reproducible_from_the_seed():
    from sorobn_b200 import examples

    a = examples.alarm(seed=7).query("Alarm", event={"John calls": True}, algorithm="gibbs", n_iterations=5000)
    b = examples.alarm(seed=7).query("Alarm", event={"John calls": True}, algorithm="gibbs", n_iterations=5000)
    c = examples.alarm(seed=8).query("Alarm", event={"John calls": True}, algorithm="gibbs", n_iterations=5000)
    pd.testing.assert_series_equal(a, b)
    assert not a.equals(c)


def test_one_chain_per_evidence_row_matches_exact_batch():
    """query_many(algorithm="gibbs"): chains are independent per row; their frequencies track
    the exact batched posteriors on a 4x4 grid with 3 states."""
    from sorobn_b200 import BayesNet, synthetic

    spec = synthetic.grid(4, 4, 3, seed=11)
    bn = synthetic.load(spec, BayesNet, seed=3)
    query = ("g0303",)
    evidence = ["g0001", "g0102", "g0210", "g0301"]
    evidence = [e for e in evidence if e in spec.nodes] or spec.nodes[1:5]
    B = 512
    events = synthetic.random_events(spec, evidence, B, seed=5)
    exact = bn.query_many(*query, events=events).to_numpy()
    gibbs = bn.query_many(*query, events=events, algorithm="gibbs", n_iterations=120_000).to_numpy()
    assert gibbs.shape == exact.shape
    assert np.allclose(gibbs.sum(axis=1), 1.0, atol=1e-5)
    err = np.abs(gibbs - exact)
    assert err.mean() < 0.01 and err.max() < 0.06, (err.mean(), err.max())


def test_gibbs_on_the_benchmark_grid_runs_and_tracks_exact():
    """configs[4] shape: the 100-node 5-state grid, one chain per evidence row."""
    from sorobn_b200 import workloads

    wl = workloads.grid10x10()
    bn = wl.build(seed=1)
    events = wl.events(256, seed=2, bn=bn)
    exact = bn.query_many(*wl.query, events=events).to_numpy()
    gibbs = bn.query_many(*wl.query, events=events, algorithm="gibbs", n_iterations=70 * 2000).to_numpy()
    assert np.allclose(gibbs.sum(axis=1), 1.0, atol=1e-5)
    # 70 non-event variables share the iterations: ~2000 draws of the query variable per chain
    assert np.abs(gibbs - exact).mean() < 0.03
