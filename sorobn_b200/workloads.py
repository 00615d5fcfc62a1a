"""The benchmark workloads BASELINE.json names, as reproducible objects.

A workload = a network, the query variables, the evidence variables (fixed for the
whole batch) and a seeded generator of evidence rows.  Rows are drawn by forward
sampling the network itself, so every row has positive probability.

    asia_1m      configs[1]  Asia (8 binary nodes), 1M evidence rows
    grid10x10    configs[2]  10x10 grid, 5 states/node, min-fill order, 100k queries
    dag50        configs[3]  50-node random DAG, <=4 parents, 8 states/node
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import pandas as pd

from . import examples, synthetic
from .bayes_net import BayesNet

__all__ = ["Workload", "asia_1m", "grid10x10", "dag50", "forward_sample_codes", "WORKLOADS"]


def forward_sample_codes(net, n_rows: int, seed: int = 0) -> np.ndarray:
    """Ancestral sampling of `n_rows` joint states of a CompiledNet, vectorised over
    rows.  Returns uint8 codes [n_vars, n_rows] (var ids are topological)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((len(net.names), n_rows), dtype=np.uint8)
    for v in range(len(net.names)):
        table = net.cpt[v]
        ps = net.parents[v]
        probs = table[tuple(out[p] for p in ps)] if ps else np.broadcast_to(table, (n_rows, table.shape[-1]))
        cdf = np.cumsum(probs, axis=-1)
        cdf = cdf / cdf[..., -1:]
        u = rng.random((n_rows, 1))
        out[v] = np.minimum((u > cdf).sum(axis=-1), table.shape[-1] - 1).astype(np.uint8)
    return out


@dataclass
class Workload:
    name: str
    description: str
    query: tuple
    evidence: tuple
    default_rows: int
    spec: object = None  # synthetic.NetSpec (synthetic networks)
    example: str | None = None  # name in examples.NETWORKS (textbook networks)

    def build(self, cls=BayesNet, **kwargs):
        if self.spec is not None:
            return synthetic.load(self.spec, cls, **kwargs)
        return examples.build(examples.NETWORKS[self.example], cls=cls, **kwargs)

    def codes(self, bn: BayesNet, n_rows: int, seed: int = 0) -> np.ndarray:
        """Evidence state codes uint8 [n_ev, n_rows] for this package's BayesNet."""
        net = bn._compiled
        allc = forward_sample_codes(net, n_rows, seed)
        return np.ascontiguousarray(allc[[net.index[v] for v in self.evidence]])

    def events(self, n_rows: int, seed: int = 0, bn: BayesNet | None = None) -> pd.DataFrame:
        """The same rows as a DataFrame of state values (one column per evidence var)."""
        bn = bn or self.build()
        net = bn._compiled
        codes = self.codes(bn, n_rows, seed)
        return pd.DataFrame({v: np.asarray(net.domains[net.index[v]], dtype=object)[codes[i]]
                             for i, v in enumerate(self.evidence)}).infer_objects()


def asia_1m() -> Workload:
    return Workload(
        name="asia_1m",
        description="Asia network (8 binary nodes): P(Lung cancer | Visit to Asia, Smoker, Positive X-ray, Dispnea), "
                    "1M independent evidence rows",
        query=("Lung cancer",),
        evidence=("Visit to Asia", "Smoker", "Positive X-ray", "Dispnea"),
        default_rows=1_000_000,
        example="asia",
    )


def grid10x10(n_evidence: int = 30, seed: int = 1) -> Workload:
    spec = synthetic.grid(10, 10, 5, seed=0)
    query = ("g0909",)  # bottom-right corner: all 100 nodes are its ancestors, all relevant
    rng = np.random.default_rng(seed)
    pool = [n for n in spec.nodes if n not in query]
    evidence = tuple(sorted(rng.choice(pool, size=n_evidence, replace=False).tolist()))
    return Workload(
        name="grid10x10",
        description=f"synthetic 10x10 grid, 5 states/node: P(g0909 | {n_evidence} observed nodes, seed {seed}), "
                    "min-fill elimination order, 100k independent evidence rows",
        query=query,
        evidence=evidence,
        default_rows=100_000,
        spec=spec,
    )


def dag50(n_evidence: int = 20, seed: int = 2) -> Workload:
    spec = synthetic.random_dag(50, 4, 8, seed=4, window=8)
    query = (spec.nodes[-1],)
    rng = np.random.default_rng(seed)
    pool = [n for n in spec.nodes if n not in query]
    evidence = tuple(sorted(rng.choice(pool, size=n_evidence, replace=False).tolist()))
    return Workload(
        name="dag50",
        description=f"synthetic 50-node random DAG, <=4 parents, 8 states/node: P({query[0]} | {n_evidence} observed "
                    f"nodes, seed {seed}), 1M independent evidence rows",
        query=query,
        evidence=evidence,
        default_rows=1_000_000,
        spec=spec,
    )


WORKLOADS = {"asia_1m": asia_1m, "grid10x10": grid10x10, "dag50": dag50}
