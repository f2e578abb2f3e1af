import sys; sys.path.insert(0,'.')
import numpy as np, torch
import pgmpy_b200 as px
from pgmpy_b200.engine import CompiledPlan
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan
for name in ["munin","diabetes"]:
    m = px.get_example_model(name); ev_vars, states = sample_evidence(m, 32, 8, seed=2)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    a = CompiledPlan(plan, "float64").run_host(states)
    b = CompiledPlan(plan, "float32").run_host(states)
    print(name, "fp32 finite:", np.isfinite(b).all(), "nan count", int(np.isnan(b).sum()), "max abs err", float(np.nanmax(np.abs(b-a))))
