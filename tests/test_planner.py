"""Planner + flat-program layout, checked on the CPU.

`oracle/program_interp.py` executes the serialised int32 program word by word with
numpy (the same words csrc/sbn_api.cu parses), so a pass here means the strides,
evidence gathers, slot reuse and axis orders the device will see are right; the
result must equal oracle.ve_oracle.query (bayes_net.py:739-794 restated) row by row.
"""
import numpy as np
import pytest

from oracle import program_interp, ve_oracle
from sorobn_b200 import BayesNet, examples, planner, synthetic, workloads


def run_plan(bn, query, ev_vars, codes, mode, **kw):
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in query], [net.index[e] for e in ev_vars], mode=mode, **kw)
    n = codes.shape[1] if codes.size else (1 if mode == planner.MODE_FLAT else 3)
    return plan, program_interp.run(plan.words, plan.table_blob64, codes, n_rows=n)


def oracle_rows(bn, query, ev_vars, codes):
    net = bn._compiled
    dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
    out = []
    for b in range(codes.shape[1]):
        ev = {v: net.domains[net.index[v]][codes[i, b]] for i, v in enumerate(ev_vars)}
        out.append(ve_oracle.query(dn, *query, event=ev)[1].reshape(-1))
    return np.stack(out, axis=1)


@pytest.mark.parametrize("trial", range(25))
def test_random_networks_batched_and_flat(trial):
    rng = np.random.default_rng(trial)
    n = int(rng.integers(3, 13))
    spec = synthetic.random_dag(n, 3, int(rng.integers(2, 5)), seed=trial)
    bn = synthetic.load(spec, BayesNet)
    perm = rng.permutation(n)
    nq = int(rng.integers(1, 3))
    ne = int(rng.integers(0, n - nq))
    query = [spec.nodes[i] for i in perm[:nq]]
    evs = [spec.nodes[i] for i in perm[nq:nq + ne]]
    B = 6
    events = synthetic.random_events(spec, evs, B, seed=trial)
    codes = np.stack([events[v].to_numpy().astype(np.uint8) for v in evs]) if evs else np.zeros((0, B), np.uint8)
    want = oracle_rows(bn, query, evs, codes) if evs else None
    plan, got = run_plan(bn, query, evs, codes, planner.MODE_BATCHED)
    if evs:
        assert np.allclose(got, want, rtol=1e-12, atol=0)
        for b in range(B):
            _, flat = run_plan(bn, query, evs, codes[:, b:b + 1], planner.MODE_FLAT)
            assert np.allclose(flat[:, 0], want[:, b], rtol=1e-12, atol=0)
    else:
        net = bn._compiled
        dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
        ref = ve_oracle.query(dn, *query, event={})[1].reshape(-1)
        assert np.allclose(got, ref[:, None], rtol=1e-12)


def test_fp32_interpreter_predicts_device_tolerance():
    """The device computes in fp32: emulate that rounding on the CPU and check it stays
    inside the 1e-6 relative budget on the benchmark grid (100 nodes, 70 eliminations)."""
    wl = workloads.grid10x10()
    bn = wl.build()
    net = bn._compiled
    codes = wl.codes(bn, 12, seed=9)
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    got64 = program_interp.run(plan.words, plan.table_blob64, codes)
    got32 = program_interp.run(plan.words, plan.table_blob, codes, dtype=np.float32)
    assert np.max(np.abs(got32 - got64) / got64) < 1e-6
    assert np.isfinite(got32).all()
    # and the float64 program equals the oracle run with the same elimination order
    dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
    order = [net.names[v] for v in plan.order]
    for b in range(3):
        ev = {v: net.domains[net.index[v]][codes[i, b]] for i, v in enumerate(wl.evidence)}
        want = ve_oracle.query(dn, *wl.query, event=ev, order=order)[1].reshape(-1)
        assert np.allclose(got64[:, b], want, rtol=1e-10)


def test_benchmark_grid_plan_shape():
    wl = workloads.grid10x10()
    bn = wl.build()
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    assert len(plan.order) == 100 - 1 - 30  # every node is an ancestor of the corner
    assert plan.Q == 5
    assert plan.max_factor_per_row() <= 5 ** 6
    assert plan.bytes_per_row() > 10_000  # a real streaming workload, not a toy
    # slot reuse keeps the per-row scratch far below the sum of all intermediates
    total = sum(int(np.prod(st.cards)) for st in plan.steps)
    assert plan.scratch_floats_per_row() < total


def test_many_factors_are_folded_in_chunks():
    """A root with 12 hidden children: eliminating it multiplies 13 factors, more than
    one launch fuses (MAX_IN = 8) -> product-only launches first (bayes_net.py:256
    reduces pairwise; the result is the same)."""
    kids = [f"k{i:02d}" for i in range(12)]
    bn = BayesNet(*[("root", k) for k in kids])
    rng = np.random.default_rng(0)
    import pandas as pd

    bn.P["root"] = pd.Series({0: 0.3, 1: 0.7})
    for k in kids:
        p = rng.random(2)
        bn.P[k] = pd.DataFrame({"root": [0, 0, 1, 1], k: [0, 1, 0, 1], "p": [p[0], 1 - p[0], p[1], 1 - p[1]]})
    bn.prepare()
    query, evs = kids[:2], kids[2:5]
    codes = np.array([[0, 1, 1], [1, 0, 1], [0, 0, 1]], dtype=np.uint8)
    # make every child *hidden-free*: query two, observe three, the other seven are hidden
    plan, got = run_plan(bn, query, evs, codes, planner.MODE_BATCHED)
    want = oracle_rows(bn, query, evs, codes)
    assert np.allclose(got, want, rtol=1e-12)
    assert max(len(st.inputs) for st in plan.steps) <= planner.MAX_IN
    plan4, got4 = run_plan(bn, query, evs, codes, planner.MODE_BATCHED, max_in=2)
    assert np.allclose(got4, want, rtol=1e-12)
    assert max(len(st.inputs) for st in plan4.steps) <= 2 and len(plan4.steps) > len(plan.steps)


def test_custom_elimination_order_and_errors():
    bn = examples.asia()
    net = bn._compiled
    q = [net.index["Dispnea"]]
    e = [net.index["Visit to Asia"]]
    hidden = [net.index[n] for n in ("Tuberculosis", "Smoker", "Lung cancer", "Bronchitis", "TB or cancer")]
    codes = np.array([[1]], dtype=np.uint8)
    base = None
    for order in (hidden, hidden[::-1]):
        plan = planner.build_plan(net, q, e, mode=planner.MODE_FLAT, order=order)
        got = program_interp.run(plan.words, plan.table_blob64, codes)
        base = got if base is None else base
        assert np.allclose(got, base, rtol=1e-13)
    with pytest.raises(ValueError):
        planner.build_plan(net, q, e, order=hidden[:-1])
    with pytest.raises(ValueError):  # bayes_net.py:840-841
        planner.build_plan(net, [], e)
    with pytest.raises(ValueError):  # bayes_net.py:843-845
        planner.build_plan(net, q, q)


def test_program_words_are_self_consistent():
    bn = examples.alarm()
    net = bn._compiled
    plan = planner.build_plan(net, [net.index["Burglary"]], [net.index["John calls"], net.index["Mary calls"]])
    hdr, tables, slots, steps = program_interp.parse(plan.words)
    assert hdr["version"] == planner.VERSION and hdr["n_ev"] == 2 and hdr["Q"] == 2
    assert all(off % 4 == 0 for off, _ in tables)  # 16-byte aligned tables (bulk-TMA copies)
    assert steps[-1]["out_slot"] == hdr["post_slot"]
    assert plan.words.dtype == np.int32 and plan.table_blob.dtype == np.float32
    # evidence axes carry (column, stride, cardinality)
    evs = [e for st in steps for i in st["inputs"] for e in i["ev"]]
    assert evs and all(0 <= col < 2 and card == 2 for col, _, card in evs)


def test_tables_are_shipped_as_plain_probabilities():
    """No rescaling: with entries <= 1 every intermediate factor stays <= 1 in fp32."""
    wl = workloads.grid10x10()
    bn = wl.build()
    net = bn._compiled
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    assert plan.table_blob.max() <= 1.0 and plan.table_blob.min() >= 0.0


def test_pure_sum_outs_are_folded_into_their_producer():
    """`sum_out(*variables)` (bayes_net.py:54) takes several variables: an elimination whose only
    factor is the product just built is folded into that launch.  Same answers, fewer bytes."""
    wl = workloads.grid10x10()
    bn = wl.build()
    net = bn._compiled
    q, e = [net.index[v] for v in wl.query], [net.index[v] for v in wl.evidence]
    merged = planner.build_plan(net, q, e, merge_sum_outs=True, fuse_elims=False)
    plain = planner.build_plan(net, q, e, merge_sum_outs=False, fuse_elims=False)
    assert len(merged.steps) < len(plain.steps)
    assert merged.bytes_per_row() < 0.9 * plain.bytes_per_row()
    assert merged.scratch_floats_per_row() <= plain.scratch_floats_per_row()
    assert any(len(st.ecards) > 1 for st in merged.steps) and all(len(st.ecards) <= 1 for st in plain.steps)
    assert all(len(st.ecards) <= planner.MAX_ELIM and st.cx <= planner.MAX_Z for st in merged.steps)
    codes = wl.codes(bn, 5, seed=2)
    a = program_interp.run(merged.words, merged.table_blob64, codes)
    b = program_interp.run(plain.words, plain.table_blob64, codes)
    assert np.allclose(a, b, rtol=1e-12, atol=0)
    # no launch writes a slot it reads, also after the merge moved outputs around
    for plan in (merged, plain):
        for st in plan.steps:
            assert all(not (f.is_slot and f.buf == st.out_slot) for f, _, _ in st.inputs)


def test_fused_eliminations_keep_answers_and_save_bytes():
    """Variables whose factors all sit in one bucket (plus tables over covered variables) are summed
    out by one launch (planner pass 5); `sum_out(*variables)` (bayes_net.py:54) is the reference's
    multi-variable form.  Same posteriors as one launch per variable, fewer bytes, valid slots."""
    for wl, min_saving in ((workloads.grid10x10(), 0.2), (workloads.dag50(), 0.1)):
        bn = wl.build()
        net = bn._compiled
        q, e = [net.index[v] for v in wl.query], [net.index[v] for v in wl.evidence]
        fused = planner.build_plan(net, q, e, fuse_elims=True)
        plain = planner.build_plan(net, q, e, fuse_elims=False)
        assert fused.bytes_per_row() < (1 - min_saving) * plain.bytes_per_row()
        assert any(len(st.ecards) > 1 for st in fused.steps) and all(len(st.ecards) <= 1 for st in plain.steps)
        assert all(len(st.ecards) <= planner.MAX_ELIM and st.cx <= planner.MAX_Z for st in fused.steps)
        # every hidden variable is eliminated exactly once
        elims = [v for st in fused.steps for v in st.elims]
        assert sorted(elims) == sorted(plain.order) and len(set(elims)) == len(elims)
        codes = wl.codes(bn, 4, seed=5)
        a = program_interp.run(fused.words, fused.table_blob64, codes)
        b = program_interp.run(plain.words, plain.table_blob64, codes)
        assert np.allclose(a, b, rtol=1e-12, atol=0)
        for st in fused.steps:
            assert all(not (f.is_slot and f.buf == st.out_slot) for f, _, _ in st.inputs)


def test_chain_collapses_to_few_launches():
    """A chain observed at the far end: every elimination after the first is a pure sum-out
    of the previous product only when no new CPT joins, so nothing merges there; but a
    query on the head with NO evidence eliminates the tail as re-layouts + sum-outs."""
    spec = synthetic.chain(9, 4, seed=3)
    bn = synthetic.load(spec, BayesNet)
    net = bn._compiled
    for q, ev in ((["c8"], []), (["c0"], ["c8"]), (["c4"], ["c0", "c8"])):
        plan = planner.build_plan(net, [net.index[v] for v in q], [net.index[v] for v in ev], mode=planner.MODE_BATCHED,
                                  merge_sum_outs=True)
        ref = planner.build_plan(net, [net.index[v] for v in q], [net.index[v] for v in ev], mode=planner.MODE_BATCHED,
                                 merge_sum_outs=False)
        codes = np.zeros((len(ev), 3), dtype=np.uint8)
        codes[:, 1] = 1
        codes[:, 2] = 3
        assert np.allclose(program_interp.run(plan.words, plan.table_blob64, codes, n_rows=3),
                           program_interp.run(ref.words, ref.table_blob64, codes, n_rows=3), rtol=1e-12)
        assert len(plan.steps) <= len(ref.steps)


def test_deferred_evidence_instantiation_keeps_answers_and_cuts_row_work():
    """Products of tables only are computed once as tables that keep their evidence axes
    (flat launches); only launches that touch a per-row factor stay batched."""
    for name in ("grid10x10", "asia_1m", "dag50"):
        wl = workloads.WORKLOADS[name]()
        bn = wl.build()
        net = bn._compiled
        q, e = [net.index[v] for v in wl.query], [net.index[v] for v in wl.evidence]
        lifted = planner.build_plan(net, q, e)
        direct = planner.build_plan(net, q, e, lift_evidence=False)
        nb = lambda p: sum(st.kind == planner.KIND_BATCHED for st in p.steps)
        assert nb(lifted) < nb(direct)
        assert lifted.bytes_per_row() < direct.bytes_per_row()
        # the tables that keep evidence axes stay within the shared-memory staging size
        for st in lifted.steps:
            if st.kind == planner.KIND_FLAT:
                assert int(np.prod(st.cards, dtype=np.int64)) <= planner.LIFT_MAX
        assert lifted.steps[-1].kind == planner.KIND_BATCHED  # the posterior itself is per row
        codes = wl.codes(bn, 7, seed=4)
        a = program_interp.run(lifted.words, lifted.table_blob64, codes)
        b = program_interp.run(direct.words, direct.table_blob64, codes)
        assert np.allclose(a, b, rtol=1e-12, atol=0)


def test_programs_without_query_variables_give_the_evidence_probability():
    """predict_proba path: no query variable, the normaliser is P(event)."""
    bn = examples.asia()
    net = bn._compiled
    dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
    for ev_names in (["Smoker", "Dispnea"], ["Positive X-ray", "Visit to Asia", "Bronchitis"], list(bn.nodes)):
        e = [net.index[v] for v in ev_names]
        with pytest.raises(ValueError):
            planner.build_plan(net, [], e)  # bayes_net.py:840 still holds for query()
        plan = planner.build_plan(net, [], e, allow_empty_query=True)
        assert plan.Q == 1
        rng = np.random.default_rng(0)
        codes = rng.integers(0, 2, size=(len(e), 9)).astype(np.uint8)
        post, totals = program_interp.run(plan.words, plan.table_blob64, codes, return_totals=True)
        for b in range(codes.shape[1]):
            ev = {v: net.domains[net.index[v]][codes[i, b]] for i, v in enumerate(ev_names)}
            want = ve_oracle.evidence_probability(dn, ev)
            assert abs(totals[b] - want) <= 1e-12 * max(want, 1e-300)
            assert want == 0 or post[0, b] == 1.0
    with pytest.raises(ValueError):
        planner.build_plan(net, [], [], allow_empty_query=True)


def test_planner_limits_raise_clear_errors():
    import pandas as pd

    # a variable with more than 255 states cannot be evidence (state codes are uint8)
    big = BayesNet(("A", "B"))
    big.P["A"] = pd.Series({k: 1 / 300 for k in range(300)})
    big.P["B"] = pd.DataFrame([(a, b, 0.5) for a in range(300) for b in (0, 1)], columns=["A", "B", "p"])
    big.prepare()
    net = big._compiled
    with pytest.raises(ValueError, match="uint8"):
        planner.build_plan(net, [net.index["B"]], [net.index["A"]])
    assert planner.build_plan(net, [net.index["A"]], [net.index["B"]]).Q == 300  # querying it is fine
    # duplicate and overlapping variables
    net = examples.alarm()._compiled
    a, b = net.index["Alarm"], net.index["Burglary"]
    with pytest.raises(ValueError):
        planner.build_plan(net, [a, a], [])
    with pytest.raises(ValueError):
        planner.build_plan(net, [a], [b, b])
    with pytest.raises(ValueError):
        planner.build_plan(net, [a], [a])
    # a factor wider than the kernel's axis limit
    wide = BayesNet(*[(f"p{k:02d}", "child") for k in range(planner.MAX_AXES + 1)])
    for k in range(planner.MAX_AXES + 1):
        wide.P[f"p{k:02d}"] = pd.Series({0: 0.5, 1: 0.5})
    wide.parents  # the CPT of `child` would have 2**22 rows: build the plan from a stub net instead
    stub = planner.CompiledNet(
        names=[f"p{k:02d}" for k in range(planner.MAX_AXES + 1)] + ["child"],
        domains=[[0, 1]] * (planner.MAX_AXES + 2),
        parents=[[] for _ in range(planner.MAX_AXES + 1)] + [list(range(planner.MAX_AXES + 1))],
        cpt=[np.array([0.5, 0.5])] * (planner.MAX_AXES + 1) + [np.zeros((1,))],
    )
    with pytest.raises(ValueError, match="axes"):
        planner.build_plan(stub, list(range(planner.MAX_AXES + 1)), [], mode=planner.MODE_FLAT)


def test_every_step_respects_the_kernel_limits():
    for name in ("grid10x10", "asia_1m", "dag50"):
        wl = workloads.WORKLOADS[name]()
        bn = wl.build()
        net = bn._compiled
        for mode in (planner.MODE_BATCHED, planner.MODE_FLAT):
            plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence], mode=mode)
            for st in plan.steps:
                assert 1 <= len(st.inputs) <= planner.MAX_IN
                assert len(st.cards) <= planner.MAX_AXES
                assert all(len(f.ev) <= planner.MAX_EV for f, _, _ in st.inputs)
                if mode == planner.MODE_FLAT:
                    assert st.kind == planner.KIND_FLAT
            # slot sizes cover what is written into them
            for st in plan.steps:
                assert plan.slots[st.out_slot][1] >= (int(np.prod(st.cards, dtype=np.int64)) if st.cards else 1)


def test_big_cpts_are_laid_out_for_their_consumer():
    """A launch whose tables exceed the shared-memory budget gets them re-shipped with the output
    axes >= 2 outermost (sliced staging, csrc plan_slices): same answers as the oracle, and the
    tiles of one value of the outermost axis touch one contiguous part of the table."""
    wl = workloads.dag50()
    bn = wl.build()
    net = bn._compiled
    dn = ve_oracle.dense_from_pandas(bn.P, bn.parents, bn.nodes)
    plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
    big = [(st, f, ss) for st in plan.steps if st.kind == planner.KIND_BATCHED
           for f, _, ss in st.inputs if not f.is_slot and net.cpt[plan.tables[f.buf]].size * 4 > planner.SLICE_MIN_BYTES]
    assert big
    for st, f, ss in big:
        rest = [s for s in ss[2:] if s]
        inner = [s for s in ss[:2] if s] + [s for _, s, _ in f.ev]
        if rest:  # (a table of a one- or two-axis launch has nothing to slice along)
            assert min(rest) > max(inner)  # axes >= 2 are the outermost ones
            assert rest == sorted(rest)  # ... in the significance order the tiles are walked in
    assert any(s for _, _, ss in big for s in ss[2:])
    codes = wl.codes(bn, 3, seed=31)
    got = program_interp.run(plan.words, plan.table_blob64, codes)
    order = [net.names[v] for v in plan.order]
    for b in range(3):
        ev = {v: net.domains[net.index[v]][int(codes[k, b])] for k, v in enumerate(wl.evidence)}
        want = ve_oracle.query(dn, *wl.query, event=ev, order=order)[1].reshape(-1)
        assert np.allclose(got[:, b], want, rtol=1e-10, atol=0)


@pytest.mark.parametrize("trial", range(60))
def test_merged_sum_outs_with_lifted_evidence_match_the_oracle(trial):
    """`merge_sum_outs=True` folds a pure sum-out into its producer.  When the producer is a
    deferred-evidence (lifted) step its output keeps evidence axes whose strides are not in the
    inputs' free-variable strides; the merged step must carry them (6 of 150 random
    networks returned wrong posteriors)."""
    rng = np.random.default_rng(1000 + trial)
    n = int(rng.integers(4, 9))
    spec = synthetic.random_dag(n, 3, int(rng.integers(2, 4)), seed=1000 + trial)
    bn = synthetic.load(spec, BayesNet)
    perm = rng.permutation(n)
    ne = int(rng.integers(1, 3))
    query = [spec.nodes[perm[0]]]
    evs = [spec.nodes[i] for i in perm[1:1 + ne]]
    B = 5
    events = synthetic.random_events(spec, evs, B, seed=trial)
    codes = np.stack([events[v].to_numpy().astype(np.uint8) for v in evs])
    want = oracle_rows(bn, query, evs, codes)
    for fuse in (False, True):
        _, got = run_plan(bn, query, evs, codes, planner.MODE_BATCHED, merge_sum_outs=True, lift_evidence=True,
                          fuse_elims=fuse)
        assert np.allclose(got, want, rtol=1e-12, atol=0), (trial, fuse)


def test_slots_and_order_allow_two_steps_per_launch():
    """What the engine's paired launches (csrc/sbn_pair.h) rely on: a batched step never writes the
    buffer an operand of the PREVIOUS batched step lives in (the fused launch reads that operand while
    it writes this output), and an expanding product is immediately followed by its consumer."""
    for name in ("grid10x10", "dag50", "asia_1m"):
        wl = workloads.WORKLOADS[name]()
        net = wl.build()._compiled
        plan = planner.build_plan(net, [net.index[q] for q in wl.query], [net.index[e] for e in wl.evidence])
        batched = [st for st in plan.steps if st.kind == planner.KIND_BATCHED]
        for prev, st in zip(batched, batched[1:]):
            held = {f.buf for f, _, _ in prev.inputs if f.is_slot and f.batched}
            held |= {f.buf for f, _, _ in st.inputs if f.is_slot and f.batched}
            assert st.out_slot not in held, (name, st.out_slot, held)
        if name == "grid10x10":
            sizes = [int(np.prod(st.cards)) for st in batched]
            big = sizes.index(3125)
            consumer = batched[big + 1]
            assert any(f.is_slot and f.batched and f.buf == batched[big].out_slot for f, _, _ in consumer.inputs)


@pytest.mark.parametrize("big_last,dfs", [("0", "1"), ("1", "0"), ("0", "0")])
def test_launch_order_switches_keep_the_answers(monkeypatch, big_last, dfs):
    """SOROBN_B200_BIG_LAST / SOROBN_B200_DFS only re-order the launches (any topological order of the
    step tree is valid): same posteriors as the oracle, slot guarantees intact."""
    monkeypatch.setenv("SOROBN_B200_BIG_LAST", big_last)
    monkeypatch.setenv("SOROBN_B200_DFS", dfs)
    for trial in range(40, 46):
        rng = np.random.default_rng(trial)
        n = int(rng.integers(6, 13))
        spec = synthetic.random_dag(n, 3, int(rng.integers(2, 5)), seed=trial)
        bn = synthetic.load(spec, BayesNet)
        perm = rng.permutation(n)
        query = [spec.nodes[perm[0]]]
        evs = [spec.nodes[i] for i in perm[1:1 + int(rng.integers(1, n - 2))]]
        events = synthetic.random_events(spec, evs, 5, seed=trial)
        codes = np.stack([events[v].to_numpy().astype(np.uint8) for v in evs])
        plan, got = run_plan(bn, query, evs, codes, planner.MODE_BATCHED)
        assert np.allclose(got, oracle_rows(bn, query, evs, codes), rtol=1e-12, atol=0)
        batched = [st for st in plan.steps if st.kind == planner.KIND_BATCHED]
        for prev, st in zip(batched, batched[1:]):
            held = {f.buf for f, _, _ in prev.inputs if f.is_slot and f.batched}
            assert st.out_slot not in held
