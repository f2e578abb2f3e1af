"""Golden most-probable-explanation answers from the UNMODIFIED reference.

TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_mpe
asia, cancer, sachs (joint tables small enough for the reference): forward-sampled evidence (seed 0), pgmpy
VariableElimination.map_query(variables = every unobserved variable, evidence) — the argmax of the full joint
(ExactInference.py:528-624) — and the joint probability P(x*, e) of the answer -> tests/golden/ref_mpe_small.json.
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

    out = {}
    for name, k, n_cases in (("asia", 2, 12), ("cancer", 1, 8), ("sachs", 3, 8)):
        rm = ref_model(name)
        m = px.get_example_model(name)
        ev_vars, states = sample_evidence(m, n_cases, k, seed=0)
        free = [v for v in m.nodes() if v not in ev_vars]
        cases = []
        for case in range(n_cases):
            ev = states_to_names(m, ev_vars, states[case])
            mp = RefVE(rm).map_query(free, evidence=ev, show_progress=False)
            full = dict(ev)
            full.update(mp)
            p = 1.0
            for cpd in rm.get_cpds():
                p *= float(cpd.get_value(**{v: full[v] for v in cpd.variables}))
            cases.append({"case": case, "map": {v: str(s) for v, s in mp.items()}, "joint_probability": p})
        out[name] = {"ev_vars": ev_vars, "ev_states": states.tolist(), "cases": cases}
        print(name, len(cases))
    out["reference"] = "pgmpy 1.0.0 VariableElimination.map_query(all unobserved variables), numpy backend"
    with open(os.path.join(OUT_DIR, "ref_mpe_small.json"), "w") as f:
        json.dump(out, f)


if __name__ == "__main__":
    main()
