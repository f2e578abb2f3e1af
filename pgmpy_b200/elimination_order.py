"""Elimination-order heuristics with the reference's class API (pgmpy/inference/EliminationOrder.py:11-166).

    WeightedMinFill(model).get_elimination_order(nodes=None, show_progress=True) -> list
    MinNeighbors / MinWeight / MinFill, BaseEliminationOrder.cost(node), .fill_in_edges(node)

Host-only graph code. The cost functions reproduce the reference's definitions INCLUDING its quirks (SURVEY.md
App. B.2), because `cost()` is public and tested against known values (tests/test_inference/test_elimination_order.py:
45-139): after a node is eliminated it is simply removed from both graphs — no fill-in edge is ever added — and
`MinFill` counts pairs of the node's CHILDREN in the directed graph, not of its moral neighbours. Those quirks are why
the reference's `"MinFill"` order needs 172 / 323 GiB tables on diabetes / munin; the engine's own
planner therefore uses `ProperMinFill` (fill-in edges added, weighted tie-break) — exposed here with the same API — and
`VariableElimination.query(elimination_order=<heuristic name>)` maps every name to it: the order changes posteriors
only at ~1e-16. Ties are broken by model node order (the reference's tie-break is hash-order
dependent, its own tests only assert sets).
"""
from __future__ import annotations

from itertools import combinations
from typing import Dict, Hashable, Iterable, List, Optional, Set

from . import graph as G
from .models import DiscreteBayesianNetwork, from_pgmpy


class BaseEliminationOrder:
    def __init__(self, model):
        if not isinstance(model, DiscreteBayesianNetwork):
            if type(model).__name__ != "DiscreteBayesianNetwork":
                raise ValueError("Model should be a DiscreteBayesianNetwork instance")
            model = from_pgmpy(model)
        self._nodes: List[Hashable] = list(model.nodes())
        self._rank = {v: i for i, v in enumerate(self._nodes)}
        self._card: Dict[Hashable, int] = dict(model.get_cardinality())
        parents = {n: list(model.get_parents(n)) for n in self._nodes}
        # the two graphs the reference keeps: the directed model (children lists) and its moral graph
        self._children: Dict[Hashable, List[Hashable]] = {n: list(model.get_children(n)) for n in self._nodes}
        self._moral: Dict[Hashable, Set[Hashable]] = G.moral_graph(parents)

    def cost(self, node) -> float:
        return 0

    def _remove(self, node) -> None:
        """What the reference does after choosing a node: drop it from both graphs, add nothing."""
        for nb in self._moral.pop(node, ()):
            self._moral[nb].discard(node)
        self._children.pop(node, None)
        for ch in self._children.values():
            if node in ch:
                ch.remove(node)

    def get_elimination_order(self, nodes: Optional[Iterable[Hashable]] = None, show_progress: bool = True) -> List[Hashable]:
        remaining = set(self._moral) if nodes is None else set(nodes)
        ordering = []
        while remaining:
            best = min(sorted(remaining, key=self._rank.get), key=self.cost)  # min() keeps the first of equal costs
            ordering.append(best)
            remaining.remove(best)
            self._remove(best)
        return ordering

    def fill_in_edges(self, node):
        return combinations(self._children.get(node, ()), 2)


class WeightedMinFill(BaseEliminationOrder):
    """cost = sum over pairs of moral neighbours of card(a) * card(b) (EliminationOrder.py:119-134)."""

    def cost(self, node):
        return sum(self._card[a] * self._card[b] for a, b in combinations(sorted(self._moral[node], key=self._rank.get), 2))


class MinNeighbors(BaseEliminationOrder):
    """cost = number of moral neighbours (:137-143)."""

    def cost(self, node):
        return len(self._moral[node])


class MinWeight(BaseEliminationOrder):
    """cost = product of the cardinalities of the moral neighbours (:146-157)."""

    def cost(self, node):
        w = 1
        for nb in self._moral[node]:
            w *= self._card[nb]
        return w


class MinFill(BaseEliminationOrder):
    """cost = number of pairs among the node's children in the directed graph (:160-166 via fill_in_edges :107-116)."""

    def cost(self, node):
        return len(list(self.fill_in_edges(node)))


class ProperMinFill(BaseEliminationOrder):
    """The planner's order (pgmpy_b200.graph.min_fill_order): fewest fill-in edges on the CURRENT moral graph, ties by
    the weight of the clique the elimination creates, fill-in edges added after every elimination."""

    def cost(self, node):
        return G._fill_count(self._moral, node)

    def get_elimination_order(self, nodes=None, show_progress=True):
        keep = [] if nodes is None else [v for v in self._moral if v not in set(nodes)]
        order, _ = G.min_fill_order({v: set(nb) for v, nb in self._moral.items()}, self._card, keep=keep, rank=self._rank)
        return order


HEURISTICS = {"weightedminfill": WeightedMinFill, "minneighbors": MinNeighbors, "minweight": MinWeight, "minfill": MinFill}
