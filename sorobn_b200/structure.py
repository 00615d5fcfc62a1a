"""Structure learning helper: the Chow-Liu tree (reference: /root/reference/sorobn/structure.py).

Host-side (pandas / numpy), like the reference: mutual information of every pair of columns,
maximum spanning tree over those weights, edges oriented away from a root.  The result feeds
`BayesNet(*edges).fit(X)`, whose queries then run on the GPU.

Off the hot path (kept from round 1; nothing here touches the device).  One deviation from the
reference: its Kruskal loop stops as soon as every vertex has a neighbour (structure.py:33-41), so on
some data it returns a FOREST; this one always completes the spanning tree.  On the reference's own
example (tests/golden/chow_liu.json) the edge lists coincide; where they would not, this function has
one extra edge per remaining component.
"""
from __future__ import annotations

import itertools

import numpy as np
import pandas as pd

__all__ = ["chow_liu", "mutual_info"]


def mutual_info(puv: pd.Series, pu: pd.Series, pv: pd.Series) -> float:
    """I(u; v) from the joint `puv` (MultiIndex [u, v]) and the marginals (structure.py:59-67)."""
    u_name, v_name = puv.index.names
    mu = pu.reindex(puv.index.get_level_values(u_name)).to_numpy()
    mv = pv.reindex(puv.index.get_level_values(v_name)).to_numpy()
    joint = puv.to_numpy()
    return float((joint * np.log(joint / (mu * mv))).sum())


class _Forest:
    """Union-find with path halving and union by size (structure.py:70-98)."""

    def __init__(self, items):
        self.parent = {x: x for x in items}
        self.size = {x: 1 for x in items}

    def find(self, x):
        while self.parent[x] != x:
            self.parent[x] = self.parent[self.parent[x]]
            x = self.parent[x]
        return x

    def union(self, a, b):
        a, b = self.find(a), self.find(b)
        if a == b:
            return False
        if self.size[a] < self.size[b]:
            a, b = b, a
        self.parent[b] = a
        self.size[a] += self.size[b]
        return True


def chow_liu(X: pd.DataFrame, root=None):
    """Edges (parent, child) of the Chow-Liu tree of `X` (structure.py:9-56): the maximum
    spanning tree of the pairwise mutual informations (Kruskal), oriented away from `root`
    (default: the first column)."""
    marginals = {c: X[c].value_counts(normalize=True) for c in X.columns}
    n = len(X)
    scored = []
    for u, v in itertools.combinations(sorted(X.columns), 2):
        joint = X.groupby([u, v]).size() / n
        scored.append((mutual_info(joint, marginals[u], marginals[v]), u, v))
    # stable sort by decreasing mutual information: ties keep the (u, v) enumeration order, as
    # the reference's `sorted(..., reverse=True)` on the same keys does
    scored.sort(key=lambda t: t[0], reverse=True)

    forest = _Forest(X.columns)
    neighbours = {c: set() for c in X.columns}
    taken = 0
    for _, u, v in scored:
        if forest.union(u, v):
            neighbours[u].add(v)
            neighbours[v].add(u)
            taken += 1
            if taken == len(X.columns) - 1:
                break

    root = X.columns[0] if root is None else root
    edges, seen, stack = [], {root}, [root]
    while stack:
        node = stack.pop()
        for nb in sorted(neighbours[node] - seen, reverse=True):
            seen.add(nb)
            edges.append((node, nb))
            stack.append(nb)
    return edges
