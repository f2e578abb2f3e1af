"""Synthetic evidence batches: forward-sample full assignments, reveal k variables (P(e) > 0 guaranteed).

Uniform-random evidence is often impossible on the bnlearn models (pathfinder CPTs are 44 % zeros,
SURVEY.md §0 fact 6), so benchmark and parity evidence is drawn by ancestral sampling from the CPTs
(SURVEY.md §8d): rng = np.random.default_rng(seed); choose the observed set E once per batch with
rng.choice(sorted(nodes), k, replace=False); evidence states = the sample's values on E.
"""
from __future__ import annotations

from typing import Hashable, List, Sequence, Tuple

import numpy as np


def topological_order(model) -> List[Hashable]:
    nodes = list(model.nodes())
    indeg = {n: len(model.get_parents(n)) for n in nodes}
    ready = [n for n in nodes if indeg[n] == 0]
    order = []
    while ready:
        n = ready.pop(0)
        order.append(n)
        for c in model.get_children(n):
            indeg[c] -= 1
            if indeg[c] == 0:
                ready.append(c)
    if len(order) != len(nodes):
        raise ValueError("model graph has a cycle")
    return order


def forward_sample(model, n: int, rng: np.random.Generator) -> Tuple[List[Hashable], np.ndarray]:
    """n joint samples as state indices, int32 [n, n_nodes] in model.nodes() order."""
    nodes = list(model.nodes())
    col = {v: i for i, v in enumerate(nodes)}
    out = np.zeros((n, len(nodes)), dtype=np.int32)
    for v in topological_order(model):
        cpd = model.get_cpds(v)
        card = int(cpd.cardinality[0])
        table = cpd.values.reshape(card, -1)
        flat = np.zeros(n, dtype=np.int64)
        for p, pc in zip(cpd.variables[1:], cpd.cardinality[1:]):
            flat = flat * int(pc) + out[:, col[p]]
        probs = table[:, flat].T  # [n, card]
        cdf = np.cumsum(probs, axis=1)
        u = rng.random(n) * cdf[:, -1]
        s = (u[:, None] >= cdf).sum(axis=1)
        s = np.minimum(s, card - 1)
        # never land on a zero-probability state through rounding at a cdf plateau
        zero = probs[np.arange(n), s] <= 0
        if zero.any():
            s[zero] = np.argmax(probs[zero], axis=1)
        out[:, col[v]] = s
    return nodes, out


def sample_evidence(model, batch: int, k: int, seed: int = 0, evidence_vars: Sequence[Hashable] = None):
    """(evidence_vars, states int32 [batch, k]) for one batch with a common observed set."""
    nodes = list(model.nodes())
    if evidence_vars is None:
        # the observed set depends on (seed, k) only, not on the batch size
        names = sorted(nodes, key=str)
        pick = np.random.default_rng([seed, 0x0E]).choice(len(names), size=k, replace=False)
        evidence_vars = [names[i] for i in pick]
    rng = np.random.default_rng(seed)
    nodes, samples = forward_sample(model, batch, rng)
    col = {v: i for i, v in enumerate(nodes)}
    states = np.ascontiguousarray(samples[:, [col[v] for v in evidence_vars]], dtype=np.int32)
    return list(evidence_vars), states


def states_to_names(model, evidence_vars, states_row) -> dict:
    st = model.states
    return {v: st[v][int(s)] for v, s in zip(evidence_vars, states_row)}
