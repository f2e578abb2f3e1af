"""Golden posteriors under soft (virtual) evidence from the UNMODIFIED reference.

TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_virtual
alarm and child: forward-sampled hard evidence (seed 0, the evidence variables of tests/golden/ref_<model>.npz), plus a
random likelihood vector on each of two unobserved variables per case (pgmpy/inference/base.py:214-299: virtual
evidence = an observed binary child per soft variable, added BEFORE pruning). Stored per case: the reference's
VariableElimination.query([q], evidence, virtual_evidence) with an explicit min-fill order (VE mode, pruned) and the
exact BP-mode value (classic elimination over ALL factors of the augmented network, no pruning — what
BeliefPropagation.query computes; same recipe as oracle/make_golden_bp.py) -> tests/golden/ref_<model>_virtual.json.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.factors.discrete import TabularCPD as RefCPD
    from pgmpy.inference import VariableElimination as RefVE
    from pgmpy.utils import get_example_model as ref_model

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence, states_to_names
    from pgmpy_b200.planner import compile_ve_plan

    for name, k, n_cases in (("alarm", 5, 24), ("child", 4, 16)):
        rm = ref_model(name)
        m = px.get_example_model(name)
        card = m.get_cardinality()
        ev_vars, states = sample_evidence(m, n_cases, k, seed=0)
        free = [v for v in sorted(m.nodes(), key=str) if v not in ev_vars]
        rng = np.random.default_rng(7)
        cases = []
        for case in range(n_cases):
            ev = states_to_names(m, ev_vars, states[case])
            pick = [free[i] for i in rng.choice(len(free), 3, replace=False)]
            soft_vars, q = pick[:2], pick[2]
            like = {v: rng.uniform(0.05, 1.0, size=card[v]) for v in soft_vars}
            virt = [RefCPD(v, card[v], like[v].reshape(-1, 1), state_names={v: list(rm.states[v])}) for v in soft_vars]
            order = [v for v in sorted(m.nodes(), key=str) if v != q and v not in ev]
            # VE mode: the reference's own pruning runs on the augmented model; names it pruned away are filtered.
            # (query() re-initialises the inference object on the augmented network — inference/base.py:299 — so every
            # case gets a fresh object, and rve.model afterwards IS the augmented network.)
            rve = RefVE(rm)
            res = rve.query([q], evidence=ev, virtual_evidence=virt, elimination_order=order, show_progress=False)
            # BP mode: all factors of the augmented network, no pruning (Markov-model path of the classic loop)
            bn = rve.model
            assert all("__" + v in bn.nodes() for v in soft_vars)
            aug_ev = dict(ev)
            aug_ev.update({"__" + v: 0 for v in soft_vars})
            full = RefVE(bn.to_markov_model())
            order_all = [v for v in bn.nodes() if v != q and v not in aug_ev]
            bp = full.query([q], evidence=aug_ev, elimination_order=order_all, show_progress=False)
            bpv = np.asarray(bp.values, dtype=np.float64)
            bpv = bpv / bpv.sum()
            cases.append({"case": case, "query": q, "soft_vars": soft_vars, "likelihoods": [like[v].tolist() for v in soft_vars],
                          "ve": np.asarray(res.values, dtype=np.float64).tolist(), "bp": bpv.tolist()})
        path = os.path.join(OUT_DIR, f"ref_{name}_virtual.json")
        with open(path, "w") as f:
            json.dump({"model": name, "ev_vars": ev_vars, "ev_states": states.tolist(), "cases": cases,
                       "reference": "pgmpy 1.0.0 VariableElimination.query(virtual_evidence=...), numpy backend, fp64"}, f)
        print(name, len(cases), "->", path)


if __name__ == "__main__":
    main()
