"""Drop-in boundary: REAL pgmpy objects (the unmodified reference, imported through oracle/ref_loader) handed to the
pgmpy_b200 inference classes. The adapters `models.from_pgmpy` / `junction_tree_from_pgmpy` must produce the same
tables, state names and plans as the package's own fixture path. CPU only: plans are compared, not executed."""
import numpy as np
import pytest

import pgmpy_b200 as px
from oracle.plan_exec import run_plan
from oracle.ref_loader import load_reference, reference_available
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.models import from_pgmpy, junction_tree_from_pgmpy
from pgmpy_b200.planner import JTStructure, build_junction_tree, compile_jt_plan, compile_ve_plan

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference (pgmpy) not present in this environment")


@pytest.fixture(scope="module")
def ref():
    load_reference()
    import pgmpy  # noqa: F401
    from pgmpy.utils import get_example_model

    return get_example_model


@pytest.mark.parametrize("name", ["asia", "alarm"])
def test_from_pgmpy_gives_the_fixture_model(ref, name):
    rm = ref(name)
    ours = px.get_example_model(name)
    got = from_pgmpy(rm)
    assert got.nodes() == ours.nodes()
    assert sorted(got.edges()) == sorted(ours.edges())
    assert got.states == ours.states
    assert got.get_cardinality() == ours.get_cardinality()
    for v in ours.nodes():
        a, b = got.get_cpds(v), ours.get_cpds(v)
        assert list(a.variables) == list(b.variables)
        np.testing.assert_array_equal(a.values, b.values)
        # the adapter reads the reference's own tables (same axis order: child first, parents in BIF order)
        np.testing.assert_array_equal(a.values, np.asarray(rm.get_cpds(v).values))
    got.check_model()


@pytest.mark.parametrize("name", ["asia", "alarm"])
def test_plans_from_a_pgmpy_model_equal_plans_from_the_fixture(ref, name):
    rm = ref(name)
    ours = px.get_example_model(name)
    got = from_pgmpy(rm)
    ev_vars, states = sample_evidence(ours, 4, 2, seed=5)
    q = [v for v in ours.nodes() if v not in ev_vars][0]
    for a, b in ((compile_ve_plan(got, [q], ev_vars), compile_ve_plan(ours, [q], ev_vars)),
                 (compile_jt_plan(JTStructure.from_model(got), ev_vars), compile_jt_plan(JTStructure.from_model(ours), ev_vars))):
        np.testing.assert_array_equal(a.pool, b.pool)
        np.testing.assert_array_equal(a.const_blob, b.const_blob)
        assert a.ev_vars == b.ev_vars and [s.vars for s in a.segments] == [s.vars for s in b.segments]


def test_inference_classes_accept_pgmpy_objects(ref):
    """VariableElimination / BeliefPropagation constructors take the reference's model (and JunctionTree) directly; the
    compiled plans are the ones the fixture path gives. (No GPU here: only construction + planning.)"""
    from pgmpy_b200 import planner as PL
    from pgmpy_b200.inference import BeliefPropagation, VariableElimination

    rm = ref("asia")
    ours = px.get_example_model("asia")
    ve = VariableElimination(rm)
    assert ve.model.nodes() == ours.nodes() and ve.cardinality == ours.get_cardinality()
    bp = BeliefPropagation(rm)
    assert bp.get_cliques() == JTStructure.from_model(ours).cliques
    # a pgmpy JunctionTree built from OUR cliques / potentials (how the goldens are made) round-trips through the adapter
    from pgmpy.factors.discrete import DiscreteFactor as RefDF
    from pgmpy.models import JunctionTree as RefJT

    jt = build_junction_tree(ours)
    rjt = RefJT()
    for c in jt.nodes():
        rjt.add_node(c)
    for u, v in jt.edges():
        rjt.add_edge(u, v)
    for f in jt.get_factors():
        rjt.add_factors(RefDF(list(f.variables), list(f.cardinality), f.values, state_names=f.state_names))
    back = junction_tree_from_pgmpy(rjt)
    assert sorted(back.nodes()) == sorted(jt.nodes())
    bp2 = BeliefPropagation(rjt)
    ev_vars, states = sample_evidence(ours, 3, 2, seed=1)
    p_a = PL.compile_jt_plan(bp2._jt, ev_vars)
    p_b = PL.compile_jt_plan(PL.JTStructure.from_junction_tree(jt), ev_vars)
    # clique order may differ between the two containers; the posteriors may not
    a = run_plan(p_a.pool, p_a.const_blob, states)
    b = run_plan(p_b.pool, p_b.const_blob, states)
    ca = {s.vars[0]: (s.out_offset, s.table.size) for s in p_a.segments}
    for s in p_b.segments:
        o, n = ca[s.vars[0]]
        np.testing.assert_allclose(a[:, o : o + n], b[:, s.out_offset : s.out_offset + s.table.size], rtol=1e-13, atol=0)


def test_config_seam_mirrors_pgmpy_config(ref):
    """pgmpy_b200.config has pgmpy.config's methods; a new inference object takes its dtype from an explicit argument,
    then pgmpy_b200.config, then the application's pgmpy.config (global_vars.py:150-189), then float64."""
    import pgmpy

    from pgmpy_b200 import config
    from pgmpy_b200.config import default_dtype, normalize_dtype
    from pgmpy_b200.inference import VariableElimination

    for meth in ("set_backend", "get_backend", "set_dtype", "get_dtype", "set_device", "get_device", "set_show_progress", "get_show_progress"):
        assert hasattr(config, meth) and hasattr(pgmpy.config, meth)
    m = px.get_example_model("asia")
    try:
        assert default_dtype() == "float64" and VariableElimination(m).dtype == "float64"
        pgmpy.config.set_dtype("float32")  # the application's pgmpy setting is honoured
        assert VariableElimination(m).dtype == "float32"
        pgmpy.config.set_dtype(np.float64)
        assert VariableElimination(m).dtype == "float64"
        config.set_backend("b200", device="cuda:0", dtype="float32")  # our own switch wins over pgmpy's
        assert config.get_backend() == "b200" and config.device_index() == 0
        assert VariableElimination(m).dtype == "float32"
        assert VariableElimination(m, dtype="float64").dtype == "float64"  # explicit argument wins over everything
        with pytest.raises(ValueError):
            config.set_backend("numpy")
        with pytest.raises(ValueError):
            config.set_device("cpu")
        with pytest.raises(ValueError):
            normalize_dtype("float16")
    finally:
        config.set_backend("b200")
        pgmpy.config.set_dtype(None)
    import torch

    assert normalize_dtype(torch.float32) == "float32" and normalize_dtype(np.dtype("float64")) == "float64"
