"""Elimination-order heuristic classes: known costs and orders of the reference's own tests
(pgmpy/tests/test_inference/test_elimination_order.py:45-139, the five-node student network)."""
import numpy as np
import pytest

from pgmpy_b200 import DiscreteBayesianNetwork, TabularCPD
from pgmpy_b200.elimination_order import (BaseEliminationOrder, MinFill, MinNeighbors, MinWeight, ProperMinFill,
                                           WeightedMinFill)


def student():
    m = DiscreteBayesianNetwork([("diff", "grade"), ("intel", "grade"), ("intel", "sat"), ("grade", "reco")])
    rng = np.random.default_rng(0)

    def cpd(v, parents):
        vals = rng.random((2, 2 ** len(parents)))
        return TabularCPD(v, 2, vals / vals.sum(axis=0), parents or None, [2] * len(parents) or None)

    m.add_cpds(cpd("diff", []), cpd("intel", []), cpd("grade", ["diff", "intel"]), cpd("sat", ["intel"]), cpd("reco", ["grade"]))
    return m


def test_base_costs_are_zero_and_fill_in_edges():
    e = BaseEliminationOrder(student())
    assert all(e.cost(v) == 0 for v in ("diff", "sat", "reco", "grade", "intel"))
    assert list(e.fill_in_edges("diff")) == []  # :40-41


def test_weighted_min_fill():  # :44-65
    e = WeightedMinFill(student())
    assert {v: e.cost(v) for v in ("diff", "sat", "reco", "grade", "intel")} == {"diff": 4, "sat": 0, "reco": 0, "grade": 12, "intel": 12}
    order = e.get_elimination_order(show_progress=False)
    assert set(order[:2]) == {"sat", "reco"} and set(order[2:]) == {"grade", "intel", "diff"}
    assert WeightedMinFill(student()).get_elimination_order(nodes=["diff", "grade", "sat"], show_progress=False) == ["sat", "diff", "grade"]


def test_min_neighbors():  # :68-88
    e = MinNeighbors(student())
    assert (e.cost("grade"), e.cost("reco"), e.cost("intel")) == (3, 1, 3)
    order = e.get_elimination_order(show_progress=False)
    assert set(order[:2]) == {"sat", "reco"} and set(order[2:]) == {"diff", "grade", "intel"}
    assert MinNeighbors(student()).get_elimination_order(nodes=["diff", "grade", "sat"], show_progress=False) == ["sat", "diff", "grade"]


def test_min_weight():  # :91-112
    e = MinWeight(student())
    assert (e.cost("diff"), e.cost("intel"), e.cost("reco")) == (4, 8, 2)
    order = e.get_elimination_order(show_progress=False)
    assert set(order[:2]) == {"sat", "reco"} and set(order[2:]) == {"diff", "intel", "grade"}
    assert MinWeight(student()).get_elimination_order(nodes=["diff", "grade", "sat"], show_progress=False) == ["sat", "diff", "grade"]


def test_min_fill_counts_children_pairs_like_the_reference():  # :115-139
    e = MinFill(student())
    assert (e.cost("diff"), e.cost("intel"), e.cost("sat")) == (0, 1, 0)
    assert set(e.get_elimination_order(show_progress=False)) == {"diff", "grade", "sat", "reco", "intel"}
    assert set(MinFill(student()).get_elimination_order(nodes=["diff", "grade", "intel"], show_progress=False)) == {"diff", "grade", "intel"}


def test_proper_min_fill_is_the_planner_order_and_rejects_other_models():
    order = ProperMinFill(student()).get_elimination_order(nodes=["diff", "grade", "sat"])
    assert sorted(order) == ["diff", "grade", "sat"]
    with pytest.raises(ValueError):
        MinFill("not a model")
