"""Golden values for the max-product queries from the UNMODIFIED reference.

TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_map
For alarm and child: forward-sampled evidence (seed 0, same evidence variables as tests/golden/ref_<model>.npz),
pgmpy VariableElimination.map_query(variables, evidence) and .max_marginal(variables, evidence) with an explicit
min-fill elimination order -> tests/golden/ref_<model>_map.json.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.inference import VariableElimination as RefVE
    from pgmpy.utils import get_example_model as ref_model

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence, states_to_names
    from pgmpy_b200.planner import compile_ve_plan

    for name, k in (("alarm", 5), ("child", 4)):
        rm = ref_model(name)
        m = px.get_example_model(name)
        ev_vars, states = sample_evidence(m, 12, k, seed=0)
        free = [v for v in sorted(m.nodes(), key=str) if v not in ev_vars]
        rng = np.random.default_rng(99)
        rve = RefVE(rm)
        cases = []
        for case in range(12):
            ev = states_to_names(m, ev_vars, states[case])
            variables = [free[i] for i in sorted(rng.choice(len(free), 3, replace=False))]
            order = list(compile_ve_plan(m, variables, ev_vars).meta["order"])
            mp = rve.map_query(variables, evidence=ev, elimination_order=order, show_progress=False)
            mm = rve.max_marginal(variables, evidence=ev, elimination_order=order, show_progress=False)
            cases.append({"case": case, "variables": variables, "map": {v: str(s) for v, s in mp.items()}, "max_marginal": float(mm)})
        path = os.path.join(OUT_DIR, f"ref_{name}_map.json")
        with open(path, "w") as f:
            json.dump({"model": name, "ev_vars": ev_vars, "ev_states": states.tolist(), "cases": cases,
                       "reference": "pgmpy 1.0.0 VariableElimination.map_query / max_marginal, numpy backend"}, f, indent=1)
        print(name, len(cases), "->", path)


if __name__ == "__main__":
    main()
