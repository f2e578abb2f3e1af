"""Host logic: graph algorithms, junction trees, plan compilation; plans are evaluated with the numpy plan
interpreter (oracle/plan_exec.py) and with the host build of the device element function (tests/hostsim)."""
import numpy as np
import pytest

import pgmpy_b200 as px
from oracle import pgm_oracle as O
from oracle.plan_exec import parse, run_plan
from pgmpy_b200 import graph as G
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.plan import MAX_OPS
from pgmpy_b200.planner import JTStructure, compile_jt_plan, compile_ve_plan, evidence_to_states

from helpers import (SIX_NODE_ANSWERS, SNOW_ANSWERS, BP_QUERY_REFERENCE_RESIDUAL, golden_bp_models, load_golden_bp, golden_models, hostsim_micro_run, hostsim_run, load_golden,
                     rel_err, six_node_net, snow_net)

ALL_MODELS = ["asia", "cancer", "sachs", "child", "alarm", "hepar2", "win95pts", "pathfinder", "munin", "diabetes"]
# (cliques, sum clique entries, sum sepset entries) of a correct min-fill tree, SURVEY.md Appendix A
JT_STATS = {"alarm": (27, 1038, 238), "hepar2": (58, 2617, 688), "win95pts": (50, 2684, 878)}


@pytest.mark.parametrize("name", ALL_MODELS)
def test_junction_tree_is_valid(name):
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    assert G.check_running_intersection(jt.cliques, jt.edges)
    assert set(v for c in jt.cliques for v in c) == set(m.nodes())
    for cpd in m.get_cpds():  # family preservation
        assert any(set(cpd.variables) <= set(c) for c in jt.cliques)
    if name in JT_STATS:
        size = lambda vs: int(np.prod([jt.card[v] for v in vs]))
        seps = sum(size(set(jt.cliques[a]) & set(jt.cliques[b])) for a, b in jt.edges)
        assert (len(jt.cliques), sum(size(c) for c in jt.cliques), seps) == JT_STATS[name]


def test_pruning_matches_oracle_kept_set():
    m = px.get_example_model("alarm")
    net = O.Net(m)
    rng = np.random.default_rng(3)
    nodes = sorted(m.nodes())
    parents = {n: m.get_parents(n) for n in nodes}
    children = {n: m.get_children(n) for n in nodes}
    for _ in range(40):
        E = list(rng.choice(nodes, 5, replace=False))
        q = [v for v in nodes if v not in E][int(rng.integers(0, 32))]
        kept, _ = O.prune(net, [q], {v: 0 for v in E})
        assert G.prune_nodes(parents, children, [q], E) == kept


def test_known_answers_through_plans():
    m = six_node_net()
    for variables, evidence, want in SIX_NODE_ANSWERS:
        plan = compile_ve_plan(m, variables, list(evidence))
        es = evidence_to_states(m.states, plan.ev_vars, [evidence])
        got = run_plan(plan.pool, plan.const_blob, es)[0]
        np.testing.assert_allclose(got, np.asarray(want).reshape(-1), atol=1e-8)
    m = snow_net()
    for variables, evidence, want in SNOW_ANSWERS:
        plan = compile_ve_plan(m, variables, list(evidence))
        es = evidence_to_states(m.states, plan.ev_vars, [evidence])
        np.testing.assert_allclose(run_plan(plan.pool, plan.const_blob, es)[0], want, atol=1e-6)
        with pytest.raises(KeyError):
            evidence_to_states(m.states, plan.ev_vars, [{"Traffic": "fast"}])


@pytest.mark.parametrize("name", golden_models())
def test_ve_plans_match_reference_golden(name):
    g = load_golden(name)
    m = px.get_example_model(name)
    limit = {"munin": 8, "diabetes": 3}.get(name, 48)
    worst = 0.0
    for case, q, want in g["ve"][:limit]:
        plan = compile_ve_plan(m, [q], g["ev_vars"])
        got = run_plan(plan.pool, plan.const_blob, g["ev_states"][case : case + 1])[0]
        worst = max(worst, rel_err(got, want))
    assert worst <= 1e-12, worst


def _jt_distributes(name):
    return ["auto"] if name in ("munin", "diabetes") else ["auto", "ss", "belief", "divide"]


@pytest.mark.parametrize("name,distribute", [(n, d) for n in golden_bp_models() for d in _jt_distributes(n)])
def test_jt_plans_match_reference_golden(name, distribute):
    """Junction-tree plans (numpy plan interpreter) vs BP-mode posteriors of the unmodified reference at a FIXED 1e-12
    (refbp_* goldens: the reference's exact classic VE over all factors, no calibration loop)."""
    g = load_golden_bp(name)
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    plan = compile_jt_plan(jt, g["ev_vars"], distribute=distribute)
    n_cases = {"munin": 2, "diabetes": 2, "pathfinder": 16}.get(name, 64)  # CPU time; the GPU tests run every case
    out = run_plan(plan.pool, plan.const_blob, g["ev_states"][:n_cases])
    col = {s.vars[0]: (s.out_offset, s.table.size) for s in plan.segments}
    worst, seen = 0.0, 0
    for case, q, want in g["items"]:
        if case < n_cases:
            o, n = col[q]
            worst = max(worst, rel_err(out[case, o : o + n], want))
            seen += 1
    assert seen > 0 and worst <= 1e-12, (seen, worst)


@pytest.mark.parametrize("name", [n for n in golden_models() if n not in ("sachs", "munin", "diabetes")])
def test_jt_plans_vs_reference_bp_query(name):
    """Secondary: the reference's BeliefPropagation.query, held to its own allclose stopping rule (fixed tolerance)."""
    g = load_golden(name)
    m = px.get_example_model(name)
    plan = compile_jt_plan(JTStructure.from_model(m), g["ev_vars"])
    out = run_plan(plan.pool, plan.const_blob, g["ev_states"])
    col = {s.vars[0]: (s.out_offset, s.table.size) for s in plan.segments}
    for case, q, want in g["bp"]:
        o, n = col[q]
        assert rel_err(out[case, o : o + n], want) <= BP_QUERY_REFERENCE_RESIDUAL, q


@pytest.mark.parametrize("name", ["asia", "alarm", "hepar2", "win95pts", "pathfinder"])
def test_hostsim_element_function_matches_plan_interpreter(name):
    """Same packed plan through the g++ build of pgx_step.cuh::contract_elem and through numpy."""
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    k = 2 if name == "asia" else 5
    ev_vars, states = sample_evidence(m, 37 if name != "pathfinder" else 3, k, seed=4)
    for distribute in ("auto", "divide"):
        plan = compile_jt_plan(jt, ev_vars, distribute=distribute)
        want = run_plan(plan.pool, plan.const_blob, states)
        for use_run in (True, False):  # contract_run (stepwise kernel) and contract_elem (generic fused kernel)
            got = hostsim_run(plan, states, use_run=use_run)
            assert rel_err(got, want) <= 1e-13
    free = [v for v in sorted(m.nodes()) if v not in ev_vars]
    plan = compile_ve_plan(m, free[:2], ev_vars, joint=True)
    assert rel_err(hostsim_run(plan, states), run_plan(plan.pool, plan.const_blob, states)) <= 1e-13
    plan = compile_ve_plan(m, free[:3], ev_vars, joint=False)
    assert rel_err(hostsim_run(plan, states), run_plan(plan.pool, plan.const_blob, states)) <= 1e-13


def test_hostsim_fp32_within_tolerance():
    m = px.get_example_model("alarm")
    jt = JTStructure.from_model(m)
    ev_vars, states = sample_evidence(m, 64, 5, seed=5)
    plan = compile_jt_plan(jt, ev_vars)
    want = run_plan(plan.pool, plan.const_blob, states)
    got = hostsim_run(plan, states, dtype=np.float32)
    assert np.max(np.abs(got - want)) <= 1e-5


def test_wide_products_are_split():
    """pathfinder's hub clique has 57 neighbours; no step may exceed MAX_OPS operands."""
    m = px.get_example_model("pathfinder")
    jt = JTStructure.from_model(m)
    plan = compile_jt_plan(jt, [])
    _, steps, _ = parse(plan.pool)
    assert max(s["K"] for s in steps) <= MAX_OPS
    assert max(len(nb) for nb in jt.nb) > MAX_OPS


def test_workspace_liveness_no_overlap():
    """A step's output never overlaps a table that is still live (checked by brute force on alarm)."""
    m = px.get_example_model("alarm")
    plan = compile_jt_plan(JTStructure.from_model(m), ["CVP", "HR"])
    live_until = {}
    for i, st in enumerate(plan.steps):
        for t, _ in st.operands:
            if t.kind == 1:
                live_until[t.tid] = i
    for s in plan.segments:
        live_until[s.table.tid] = len(plan.steps)
    born = {st.out.tid: (i, st.out) for i, st in enumerate(plan.steps)}
    for tid, (i, t) in born.items():
        for tid2, (j, t2) in born.items():
            if tid2 == tid or j >= i:
                continue
            if live_until.get(tid2, j) >= i:  # t2 still live when t is written
                assert t.offset + t.size <= t2.offset or t2.offset + t2.size <= t.offset


def test_evidence_sampler_is_deterministic_and_possible():
    m = px.get_example_model("pathfinder")
    ev1, s1 = sample_evidence(m, 16, 8, seed=0)
    ev2, s2 = sample_evidence(m, 16, 8, seed=0)
    assert ev1 == ev2 and np.array_equal(s1, s2)
    plan = compile_jt_plan(JTStructure.from_model(m), ev1)
    out = run_plan(plan.pool, plan.const_blob, s1)
    assert np.isfinite(out).all()  # forward-sampled evidence always has P(e) > 0


@pytest.mark.parametrize("name", ["asia", "alarm", "hepar2", "win95pts", "pathfinder"])
def test_microprogram_tables_match_plan_interpreter(name):
    """Offset tables / level partition / column map built by pgx_fused.cuh::build_micro, walked on the CPU."""
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    ev_vars, states = sample_evidence(m, 9 if name != "pathfinder" else 2, 2 if name == "asia" else 5, seed=6)
    for distribute in (("ss", "auto", "divide") if name != "pathfinder" else ("belief",)):
        plan = compile_jt_plan(jt, ev_vars, distribute=distribute)
        got, n_levels = hostsim_micro_run(plan, states)
        assert n_levels == max(st.level for st in plan.steps) + 1
        assert rel_err(got, run_plan(plan.pool, plan.const_blob, states)) <= 1e-13
    free = [v for v in sorted(m.nodes()) if v not in ev_vars]
    plan = compile_ve_plan(m, free[:3], ev_vars, joint=False)
    got, _ = hostsim_micro_run(plan, states)
    assert rel_err(got, run_plan(plan.pool, plan.const_blob, states)) <= 1e-13


@pytest.mark.parametrize("name", ["asia", "alarm", "win95pts"])
def test_multi_query_ve_plan_matches_reference_golden(name):
    """All VE-mode single-variable posteriors of one evidence signature in ONE plan (each query keeps its own pruning)."""
    from pgmpy_b200.planner import compile_ve_multi_plan

    g = load_golden(name)
    m = px.get_example_model(name)
    free = [v for v in m.nodes() if v not in g["ev_vars"]]
    plan = compile_ve_multi_plan(m, [[v] for v in free], g["ev_vars"])
    out = run_plan(plan.pool, plan.const_blob, g["ev_states"])
    col = {s.vars[0]: (s.out_offset, s.table.size) for s in plan.segments}
    assert max(rel_err(out[case, col[q][0] : col[q][0] + col[q][1]], want) for case, q, want in g["ve"]) <= 1e-12


def test_induced_graph_and_width_known_answers():
    """pgmpy/tests/test_inference/test_ExactInference.py:341-364 (host graph code; no GPU involved).
    The constructor only needs the model, so it is built without touching the engine."""
    from pgmpy_b200.inference import VariableElimination

    ve = VariableElimination.__new__(VariableElimination)
    ve.model = six_node_net()
    ve.variables = set(ve.model.nodes())
    g = ve.induced_graph(["G", "Q", "A", "J", "L", "R"])
    assert sorted(sorted(e) for e in g.edges()) == [
        ["A", "J"], ["A", "R"], ["G", "J"], ["G", "L"], ["J", "L"], ["J", "Q"], ["J", "R"], ["L", "R"]]
    assert ve.induced_width(["G", "Q", "A", "J", "L", "R"]) == 2
    with pytest.raises(ValueError):
        ve.induced_graph(["G", "Q"])


def test_plan_save_load_round_trip(tmp_path):
    """A plan written to disk and read back executes identically (plan cache)."""
    from pgmpy_b200.plan import Plan

    m = px.get_example_model("alarm")
    ev_vars, states = sample_evidence(m, 7, 5, seed=8)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    path = str(tmp_path / "alarm_plan.npz")
    plan.save(path)
    back = Plan.load(path)
    assert np.array_equal(back.pool, plan.pool) and np.array_equal(back.const_blob, plan.const_blob)
    assert list(back.ev_vars) == list(plan.ev_vars) and back.out_elems == plan.out_elems
    assert [(s.vars, s.out_offset, s.table.size) for s in back.segments] == [(tuple(map(str, s.vars)), s.out_offset, s.table.size) for s in plan.segments]
    np.testing.assert_array_equal(run_plan(back.pool, back.const_blob, states), run_plan(plan.pool, plan.const_blob, states))
    np.testing.assert_array_equal(hostsim_run(back, states), hostsim_run(plan, states))


def _load_virtual_golden(name):
    import json
    import os

    with open(os.path.join(os.path.dirname(__file__), "golden", f"ref_{name}_virtual.json")) as f:
        return json.load(f)


@pytest.mark.parametrize("name", ["alarm", "child"])
def test_soft_evidence_plans_match_reference_golden(name):
    """Soft (virtual) evidence as per-evidence-set INPUT tables of the plan against posteriors
    the unmodified reference produced with virtual_evidence=[TabularCPD...] (oracle/make_golden_virtual.py): VE mode
    (its pruning sees the soft variables as parents of evidence, inference/base.py:256-299 then :154-212) and BP mode
    (junction-tree plan, dense and factorized potentials), fixed 1e-12. Cases that share a signature run as ONE batch
    with a different likelihood row per evidence set."""
    g = _load_virtual_golden(name)
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    ev_vars = g["ev_vars"]
    states = np.asarray(g["ev_states"], dtype=np.int32)
    for c in g["cases"]:
        soft_vars, q = c["soft_vars"], c["query"]
        row = np.concatenate([np.asarray(l, dtype=np.float64) for l in c["likelihoods"]])
        ve = compile_ve_plan(m, [q], ev_vars, soft_vars=soft_vars)
        assert [v[0] for v, _, _ in ve.inputs] == list(soft_vars) and ve.in_elems == row.size
        # a batch of three evidence sets with three different likelihood rows: row 1 carries the golden case
        soft = np.stack([np.ones_like(row), row, row[::-1] * 0.5 + 0.1])
        ev3 = np.stack([states[(c["case"] + 1) % len(states)], states[c["case"]], states[c["case"]]])
        got = run_plan(ve.pool, ve.const_blob, ev3, soft=soft)
        assert rel_err(got[1], np.asarray(c["ve"])) <= 1e-12
        for factorized in (False, True):
            bp = compile_jt_plan(jt, ev_vars, [q], soft_vars=soft_vars, factorized=factorized, distribute="ss")
            got = run_plan(bp.pool, bp.const_blob, ev3, soft=soft)
            assert rel_err(got[1], np.asarray(c["bp"])) <= 1e-12
        # uniform likelihoods change nothing: row 0 equals the plain BP-mode plan on that evidence set
        plain = compile_jt_plan(jt, ev_vars, [q], distribute="ss", factorized=False)
        assert rel_err(got[0], run_plan(plain.pool, plain.const_blob, ev3[:1])[0]) <= 1e-12
    with pytest.raises(ValueError):
        compile_ve_plan(m, [g["cases"][0]["query"]], ev_vars, soft_vars=[ev_vars[0]])


def test_soft_evidence_all_marginals_plan_and_save_load(tmp_path):
    """All-marginals junction-tree plan with two soft-evidence inputs: the auto-selected plan variant equals the dense
    Shafer-Shenoy one, and the inputs block survives the plan cache."""
    from pgmpy_b200.plan import Plan

    m = px.get_example_model("alarm")
    jt = JTStructure.from_model(m)
    ev_vars, states = sample_evidence(m, 6, 5, seed=3)
    soft_vars = [v for v in sorted(m.nodes()) if v not in ev_vars][:2]
    rng = np.random.default_rng(1)
    soft = rng.uniform(0.1, 1.0, size=(6, sum(m.get_cardinality()[v] for v in soft_vars)))
    a = compile_jt_plan(jt, ev_vars, soft_vars=soft_vars)
    b = compile_jt_plan(jt, ev_vars, soft_vars=soft_vars, distribute="ss", factorized=False)
    want = run_plan(b.pool, b.const_blob, states, soft=soft)
    assert rel_err(run_plan(a.pool, a.const_blob, states, soft=soft), want) <= 1e-12
    path = str(tmp_path / "plan.npz")
    a.save(path)
    z = Plan.load(path)
    assert z.in_elems == a.in_elems and [tuple(v) for v, _, _ in z.inputs] == [tuple(v) for v, _, _ in a.inputs]
    assert rel_err(run_plan(z.pool, z.const_blob, states, soft=soft), want) <= 1e-12


def _joint_probability(m, assignment):
    """P(x) of a full assignment {var: state index} as the product of the CPT entries."""
    p = 1.0
    for cpd in m.get_cpds():
        p *= float(cpd.values[tuple(int(assignment[v]) for v in cpd.variables)])
    return p


def _max_product_value(m, ev_vars, ev_row):
    """max over the unobserved variables of P(x, e): max-elimination with the oracle's factor algebra, in an order that
    has nothing to do with the junction tree (independent check of the traceback's answer)."""
    facs = []
    ev = dict(zip(ev_vars, (int(s) for s in ev_row)))
    for cpd in m.get_cpds():
        f = O.Factor(list(cpd.variables), np.asarray(cpd.values, dtype=np.float64))
        here = [(v, ev[v]) for v in f.variables if v in ev]
        facs.append(O.reduce(f, here) if here else f)
    for v in sorted((v for v in m.nodes() if v not in ev), key=str):
        touch = [f for f in facs if v in f.variables]
        rest = [f for f in facs if v not in f.variables]
        prod = O.factor_product(*touch)
        facs = rest + [O.maximize(prod, [v])]
    val = 1.0
    for f in facs:
        val *= float(np.asarray(f.values).reshape(-1)[0])
    return val


def test_mpe_plan_and_traceback_vs_reference_golden():
    """Max-product junction-tree plan + traceback (numpy restatement of k_mpe_traceback) against the reference's
    map_query over every unobserved variable (tests/golden/ref_mpe_small.json, oracle/make_golden_mpe.py): same joint
    probability to 1e-12, and the same assignment wherever the maximum is unique."""
    import json
    import os

    from oracle.plan_exec import mpe_traceback
    from pgmpy_b200.planner import compile_jt_mpe_plan

    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_mpe_small.json")) as f:
        g = json.load(f)
    for name in ("asia", "cancer", "sachs"):
        m = px.get_example_model(name)
        ev_vars = g[name]["ev_vars"]
        states = np.asarray(g[name]["ev_states"], dtype=np.int32)
        plan, trace, cols = compile_jt_mpe_plan(JTStructure.from_model(m), ev_vars)
        assert plan.out_elems == 0 and set(cols) == set(m.nodes()) - set(ev_vars)
        _, ws = run_plan(plan.pool, plan.const_blob, states, return_workspace=True)
        asg = mpe_traceback(trace, ws)
        for c in g[name]["cases"]:
            full = dict(zip(ev_vars, states[c["case"]]))
            full.update(dict(zip(cols, asg[c["case"]])))
            p = _joint_probability(m, full)
            assert abs(p - c["joint_probability"]) <= 1e-12 * c["joint_probability"]
            names = {v: str(m.states[v][int(s)]) for v, s in zip(cols, asg[c["case"]])}
            if names != c["map"]:  # only an exact tie may be resolved differently
                assert p == c["joint_probability"], (name, c["case"])


@pytest.mark.parametrize("name", ["alarm", "child"])
def test_mpe_plan_reaches_the_max_product_value(name):
    """Where the reference's full-joint argmax is infeasible (alarm: 1e15 entries): the traceback's assignment must
    have joint probability equal to max_x P(x, e) computed by an independent max-elimination."""
    from oracle.plan_exec import mpe_traceback
    from pgmpy_b200.planner import compile_jt_mpe_plan

    m = px.get_example_model(name)
    ev_vars, states = sample_evidence(m, 6, 4, seed=5)
    plan, trace, cols = compile_jt_mpe_plan(JTStructure.from_model(m), ev_vars)
    _, ws = run_plan(plan.pool, plan.const_blob, states, return_workspace=True)
    asg = mpe_traceback(trace, ws)
    for b in range(len(states)):
        full = dict(zip(ev_vars, states[b]))
        full.update(dict(zip(cols, asg[b])))
        want = _max_product_value(m, ev_vars, states[b])
        assert abs(_joint_probability(m, full) - want) <= 1e-12 * want
