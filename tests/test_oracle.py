"""The oracle (numpy restatement of the reference) against the reference's own known answers and against
golden posteriors produced by the unmodified reference (tests/golden/, oracle/make_golden.py)."""
import numpy as np
import pytest

import pgmpy_b200 as px
from oracle import pgm_oracle as O
from pgmpy_b200.planner import JTStructure

from helpers import (SIX_NODE_ANSWERS, SNOW_ANSWERS, BP_QUERY_REFERENCE_RESIDUAL, golden_bp_models, load_golden_bp, golden_models, load_golden, rel_err,
                     six_node_net, snow_net)


def test_six_node_known_answers():
    net = O.Net(six_node_net())
    for variables, evidence, want in SIX_NODE_ANSWERS:
        got = O.ve_query(net, variables, evidence)
        np.testing.assert_allclose(got.values, want, atol=1e-8)


def test_snow_known_answers_named_states():
    net = O.Net(snow_net())
    for variables, evidence, want in SNOW_ANSWERS:
        got = O.ve_query(net, variables, evidence)
        np.testing.assert_allclose(got.values, want, atol=1e-6)


def test_factor_algebra_vectors():
    """pgmpy/tests/test_factors/test_discrete/test_Factor.py: marginalize :390-426, normalize :452-466,
    reduce :508-530, product :582-648, divide :674-712."""
    phi = O.Factor(["x1", "x2", "x3"], np.arange(12, dtype=float).reshape(3, 2, 2))
    np.testing.assert_array_equal(O.marginalize(phi, ["x1"]).values.reshape(-1), [12, 15, 18, 21])
    np.testing.assert_array_equal(O.marginalize(phi, ["x1", "x2"]).values, [30, 36])
    with pytest.raises(ValueError):
        O.marginalize(phi, ["x4"])
    n = O.normalize(O.Factor(["x1", "x2", "x3"], np.arange(12, dtype=float).reshape(2, 3, 2)))
    np.testing.assert_allclose(n.values.reshape(-1), np.arange(12) / 66.0)
    r = O.reduce(phi, [("x3", 0), ("x2", 0)])
    assert r.variables == ["x1"]
    np.testing.assert_array_equal(r.values, [0, 4, 8])
    a = O.Factor(["x1", "x2"], np.arange(4, dtype=float).reshape(2, 2))
    b = O.Factor(["x3", "x4"], np.arange(4, dtype=float).reshape(2, 2))
    p = O.product(a, b)
    np.testing.assert_array_equal(p.values.reshape(-1), [0, 0, 0, 0, 0, 1, 2, 3, 0, 2, 4, 6, 0, 3, 6, 9])
    c = O.Factor(["x3", "x1"], np.arange(4, dtype=float).reshape(2, 2))  # shared variable, different axis order
    q = O.product(a, c)
    want = np.einsum("ij,ki->ijk", a.values, c.values)
    np.testing.assert_array_equal(O.reorder(q, ["x1", "x2", "x3"]), want)
    d = O.divide(O.Factor(["x1", "x2", "x3"], np.arange(1, 13, dtype=float).reshape(2, 2, 3)),
                 O.Factor(["x3", "x1"], np.arange(1, 7, dtype=float).reshape(3, 2)))
    np.testing.assert_allclose(
        d.values.reshape(-1), [1.0, 0.6666667, 0.6, 4.0, 1.6666667, 1.2, 3.5, 2.0, 1.5, 5.0, 2.75, 2.0], atol=1e-6)
    z = O.divide(O.Factor(["x1", "x2", "x3"], np.arange(1, 13, dtype=float).reshape(2, 2, 3)),
                 O.Factor(["x3"], [2.0, 0.0, 2.0]))
    assert np.isinf(z.values[:, :, 1]).all()  # x/0 stays inf
    zz = O.divide(O.Factor(["a"], [0.0, 1.0]), O.Factor(["a"], [0.0, 2.0]))
    np.testing.assert_array_equal(zz.values, [0.0, 0.5])  # 0/0 -> 0


@pytest.mark.parametrize("name", golden_models())
def test_ve_mode_matches_reference_golden(name):
    """VariableElimination.query of the unmodified reference (with its pruning) vs the oracle restatement."""
    g = load_golden(name)
    m = px.get_example_model(name)
    net = O.Net(m)
    states = m.states
    limit = {"munin": 6, "diabetes": 2, "pathfinder": 16}.get(name, 64)
    worst = 0.0
    for case, q, want in g["ve"][:limit]:
        ev = {v: states[v][int(s)] for v, s in zip(g["ev_vars"], g["ev_states"][case])}
        got = O.ve_query(net, [q], ev)
        worst = max(worst, rel_err(got.values, want))
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("name", golden_bp_models())
def test_bp_mode_closed_form_matches_reference_golden(name):
    """Oracle BP mode (all factors, no pruning) vs the reference's exact classic VE over all factors, fixed 1e-12."""
    g = load_golden_bp(name)
    m = px.get_example_model(name)
    net = O.Net(m)
    limit = {"munin": 6, "diabetes": 3, "pathfinder": 16}.get(name, 64)
    worst = 0.0
    for case, q, want in g["items"][:limit]:
        ev = {v: m.states[v][int(s)] for v, s in zip(g["ev_vars"], g["ev_states"][case])}
        got = O.ve_query(net, [q], ev, prune_model=False)
        worst = max(worst, rel_err(got.values, want))
    assert worst <= 1e-12, worst


@pytest.mark.parametrize("name", [n for n in golden_models() if n in ("asia", "cancer", "child", "alarm", "hepar2")])
def test_bp_mode_matches_reference_golden(name):
    """BeliefPropagation(<our junction tree>).query of the unmodified reference vs the oracle BP: both iterate belief
    updates until np.allclose accepts, so they agree to that stopping rule (fixed tolerance, no adjustment)."""
    g = load_golden(name)
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    bp = O.BP(jt.cliques, jt.edges, [O.Factor(c, p) for c, p in zip(jt.cliques, jt.potentials)])
    bp.calibrate()
    for case, q, want in g["bp"][:24]:
        ev_idx = {v: int(s) for v, s in zip(g["ev_vars"], g["ev_states"][case])}
        got = bp.query([q], ev_idx)
        assert rel_err(got.values, want) <= BP_QUERY_REFERENCE_RESIDUAL


def test_bp_closed_form_equals_unpruned_ve():
    """BP mode == normalise(sum prod of ALL CPTs) (no pruning); VE mode differs at ~1e-8 on
    the shipped BIFs because pruning renormalises CPT columns that sum to 1 only to ~1e-7."""
    g = load_golden("alarm")
    m = px.get_example_model("alarm")
    net = O.Net(m)
    for case, q, want in g["bp"][:12]:
        ev = {v: m.states[v][int(s)] for v, s in zip(g["ev_vars"], g["ev_states"][case])}
        got = O.ve_query(net, [q], ev, prune_model=False)
        assert rel_err(got.values, want) <= 1e-12
