"""The textbook networks the reference ships (/root/reference/sorobn/examples.py):
alarm, asia, sprinkler, grades.  Same structure and probabilities, stated as data:
for every node, its parents in CPT-column order, its states, and one probability row
per parent combination (states in the order listed).
"""
from __future__ import annotations

import itertools

import pandas as pd

from .bayes_net import BayesNet

__all__ = ["alarm", "asia", "sprinkler", "grades", "NETWORKS"]

T, F = True, False

# node: (parents, states, {parent values: probabilities of `states`})
_ALARM = {
    "Burglary": ((), (T, F), {(): (0.001, 0.999)}),
    "Earthquake": ((), (T, F), {(): (0.002, 0.998)}),
    "Alarm": (("Burglary", "Earthquake"), (T, F), {
        (T, T): (0.95, 0.05), (T, F): (0.94, 0.06), (F, T): (0.29, 0.71), (F, F): (0.001, 0.999)}),
    "John calls": (("Alarm",), (T, F), {(T,): (0.9, 0.1), (F,): (0.05, 0.95)}),
    "Mary calls": (("Alarm",), (T, F), {(T,): (0.7, 0.3), (F,): (0.01, 0.99)}),
}

_ASIA = {
    "Visit to Asia": ((), (T, F), {(): (0.01, 0.99)}),
    "Tuberculosis": (("Visit to Asia",), (T, F), {(T,): (0.05, 0.95), (F,): (0.01, 0.99)}),
    "Smoker": ((), (T, F), {(): (0.5, 0.5)}),
    "Lung cancer": (("Smoker",), (T, F), {(T,): (0.1, 0.9), (F,): (0.01, 0.99)}),
    "Bronchitis": (("Smoker",), (T, F), {(T,): (0.6, 0.4), (F,): (0.3, 0.7)}),
    "TB or cancer": (("Lung cancer", "Tuberculosis"), (T, F), {
        (T, T): (1, 0), (T, F): (1, 0), (F, T): (1, 0), (F, F): (0, 1)}),
    "Positive X-ray": (("TB or cancer",), (T, F), {(T,): (0.98, 0.02), (F,): (0.05, 0.95)}),
    "Dispnea": (("Bronchitis", "TB or cancer"), (T, F), {
        (T, T): (0.9, 0.1), (T, F): (0.7, 0.3), (F, T): (0.8, 0.2), (F, F): (0.1, 0.9)}),
}

_SPRINKLER = {
    "Cloudy": ((), (F, T), {(): (0.5, 0.5)}),
    "Sprinkler": (("Cloudy",), (T, F), {(T,): (0.1, 0.9), (F,): (0.5, 0.5)}),
    "Rain": (("Cloudy",), (T, F), {(T,): (0.8, 0.2), (F,): (0.2, 0.8)}),
    "Wet grass": (("Rain", "Sprinkler"), (T, F), {
        (T, T): (0.99, 0.01), (T, F): (0.9, 0.1), (F, T): (0.9, 0.1), (F, F): (0, 1)}),
}

_GRADES = {
    "Difficulty": ((), ("Easy", "Hard"), {(): (0.6, 0.4)}),
    "Intelligence": ((), ("Average", "Smart"), {(): (0.7, 0.3)}),
    "Grade": (("Difficulty", "Intelligence"), ("A", "B", "C"), {
        ("Easy", "Average"): (0.3, 0.4, 0.3), ("Easy", "Smart"): (0.9, 0.08, 0.02),
        ("Hard", "Average"): (0.05, 0.25, 0.7), ("Hard", "Smart"): (0.5, 0.3, 0.2)}),
    "SAT": (("Intelligence",), ("Failure", "Success"), {("Average",): (0.95, 0.05), ("Smart",): (0.2, 0.8)}),
    "Letter": (("Grade",), ("Weak", "Strong"), {("A",): (0.1, 0.9), ("B",): (0.4, 0.6), ("C",): (0.99, 0.01)}),
}

NETWORKS = {"alarm": _ALARM, "asia": _ASIA, "sprinkler": _SPRINKLER, "grades": _GRADES}


def tables(spec: dict) -> dict:
    """node -> pandas object for `BayesNet.P` (Series for roots, DataFrame with a 'p'
    column otherwise, the two input forms `prepare()` accepts)."""
    out = {}
    for node, (parents, states, rows) in spec.items():
        if not parents:
            out[node] = pd.Series(dict(zip(states, rows[()])))
            continue
        records = []
        for combo, probs in rows.items():
            for state, p in zip(states, probs):
                records.append((*combo, state, p))
        out[node] = pd.DataFrame(records, columns=[*parents, node, "p"])
    return out


def build(spec: dict, cls=BayesNet, **kwargs):
    """Instantiate `cls` (this package's BayesNet by default; the reference's class in
    oracle/gen_golden.py) from one of the specs above."""
    edges = [(p, node) for node, (parents, _, _) in spec.items() for p in parents]
    lone = [node for node, (parents, _, _) in spec.items()
            if not parents and not any(node in ps for ps, _, _ in spec.values())]
    bn = cls(*edges, *lone, **kwargs)
    for node, table in tables(spec).items():
        bn.P[node] = table
    bn.prepare()
    return bn


def alarm(**kwargs) -> BayesNet:
    """Judea Pearl's burglary/earthquake alarm network (5 binary nodes)."""
    return build(_ALARM, **kwargs)


def asia(**kwargs) -> BayesNet:
    """Lauritzen & Spiegelhalter's Asia chest-clinic network (8 binary nodes)."""
    return build(_ASIA, **kwargs)


def sprinkler(**kwargs) -> BayesNet:
    """AIMA figure 14.12(a): cloudy / sprinkler / rain / wet grass."""
    return build(_SPRINKLER, **kwargs)


def grades(**kwargs) -> BayesNet:
    """Koller & Friedman's student network."""
    return build(_GRADES, **kwargs)


def all_events(bn, evidence_vars):
    """Every joint assignment of `evidence_vars` (used by tests and golden vectors)."""
    net = bn._compiled
    doms = [net.domains[net.index[v]] for v in evidence_vars]
    return [dict(zip(evidence_vars, combo)) for combo in itertools.product(*doms)]
