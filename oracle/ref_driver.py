"""Drive the REAL reference (oracle/_ref) on this repo's workloads.

TEST / BENCH INFRASTRUCTURE ONLY: imported by tests/, oracle/gen_golden.py and the CPU legs of
bench.py, never by sorobn_b200/.

`ordered_query` is the reference's exact inference with a GIVEN elimination order.
`BayesNet._variable_elimination` (/root/reference/sorobn/bayes_net.py:739-794) eliminates the
hidden nodes in Python-set iteration order, which on the 10x10 benchmark grid builds factors
that do not fit in memory (the process is OOM-killed).  BASELINE.json asks for the min-fill
order, so this walks the same loop with the order fixed and calls the reference's OWN operators
for everything numeric: `pointwise_mul` (bayes_net.py:253-256) and `.cdt.sum_out`
(bayes_net.py:54-103).
"""
from __future__ import annotations


def ordered_query(ref, bn, query, event, order):
    """-> pandas Series, the reference's answer (bayes_net.py:788-794, :872-875)."""
    pm = ref.bayes_net.pointwise_mul
    relevant = {*query, *event}
    for node in list(relevant):
        relevant |= bn.ancestors(node)
    hidden = relevant - {*query, *event}
    assert set(order) == hidden
    factors = []
    for node in relevant:
        factor = bn.P[node].copy()
        for var, val in event.items():
            if var in factor.index.names:
                factor = factor[factor.index.get_level_values(var) == val]
        factors.append(factor)
    for node in order:
        prod = pm(factors.pop(i) for i in reversed(range(len(factors))) if node in factors[i].index.names)
        factors.append(prod.cdt.sum_out(node))
    posterior = pm(factors)
    posterior = posterior / posterior.sum()
    posterior.index = posterior.index.droplevel(list(set(posterior.index.names) - set(query)))
    return posterior.rename(f"P({', '.join(query)})").sort_index()


def build_workload(ref, wl):
    """The workload's network as a reference `BayesNet` (prepared)."""
    return wl.build(cls=ref.BayesNet)
