"""Generate golden posteriors by running the UNMODIFIED reference (pgmpy at /root/reference).

TEST INFRASTRUCTURE; build container only:   python -m oracle.make_golden [model ...]

For each model: forward-sampled evidence (pgmpy_b200.evidence.sample_evidence, seed 0), then
  * VE mode  pgmpy VariableElimination(model).query([q], evidence, elimination_order=<explicit min-fill list>)
             — fully in-tree numpy arithmetic incl. pgmpy's pruning (SURVEY.md §8c oracle protocol);
  * BP mode  pgmpy BeliefPropagation(<our min-fill JunctionTree as a pgmpy JunctionTree>).query([q], evidence)
             — needs the opt_einsum stand-in of oracle/shims (results differ between contraction orders
             only at ~1e-16).
Writes tests/golden/ref_<model>.npz: evidence variables/states, query list, concatenated posteriors.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# model -> (k evidence vars, n VE cases, VE queries per case (None = all), n BP cases, BP queries per case)
SPEC = {
    "asia": (2, 8, None, 8, None),
    "cancer": (1, 4, None, 4, None),
    "sachs": (3, 6, None, 4, None),
    "child": (4, 6, None, 3, None),
    "alarm": (5, 16, None, 6, None),
    "hepar2": (8, 6, None, 2, 24),
    "win95pts": (8, 6, None, 2, 24),
    "pathfinder": (8, 3, 16, 1, 6),
    "munin": (8, 2, 10, 0, 0),
    "diabetes": (8, 1, 6, 0, 0),
}


def main(names):
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.factors.discrete import DiscreteFactor as RefDF
    from pgmpy.inference import BeliefPropagation as RefBP
    from pgmpy.inference import VariableElimination as RefVE
    from pgmpy.models import JunctionTree as RefJT
    from pgmpy.utils import get_example_model as ref_model

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence, states_to_names
    from pgmpy_b200.planner import JTStructure, compile_ve_plan

    for name in names:
        k, n_ve, q_ve, n_bp, q_bp = SPEC[name]
        t0 = time.time()
        rm = ref_model(name)
        m = px.get_example_model(name)
        ev_vars, states = sample_evidence(m, max(n_ve, n_bp, 1), k, seed=0)
        free = [v for v in sorted(m.nodes(), key=str) if v not in ev_vars]
        rng = np.random.default_rng(12345)
        rve = RefVE(rm)
        ve_q, ve_vals = [], []
        for case in range(n_ve):
            ev = states_to_names(m, ev_vars, states[case])
            qs = free if q_ve is None else [free[i] for i in sorted(rng.choice(len(free), q_ve, replace=False))]
            for q in qs:
                order = list(compile_ve_plan(m, [q], ev_vars).meta["order"])
                res = rve.query([q], evidence=ev, elimination_order=order, show_progress=False)
                ve_q.append([case, q])
                ve_vals.append(np.asarray(res.values, dtype=np.float64).reshape(-1))
        bp_q, bp_vals = [], []
        if n_bp:
            jt = JTStructure.from_model(m)
            if any(not (set(jt.cliques[a]) & set(jt.cliques[b])) for a, b in jt.edges):
                n_bp = 0  # disconnected network: pgmpy's ClusterGraph.add_edge refuses empty sepsets
        if n_bp:
            rjt = RefJT()
            for c in jt.cliques:
                rjt.add_node(c)
            for a, b in jt.edges:
                rjt.add_edge(jt.cliques[a], jt.cliques[b])
            for c, p in zip(jt.cliques, jt.potentials):
                rjt.add_factors(RefDF(list(c), [jt.card[v] for v in c], p, state_names={v: m.states[v] for v in c}))
            rbp = RefBP(rjt)
            for case in range(n_bp):
                ev = states_to_names(m, ev_vars, states[case])
                qs = free if q_bp is None else [free[i] for i in sorted(rng.choice(len(free), q_bp, replace=False))]
                for q in qs:
                    with np.errstate(all="ignore"):
                        res = rbp.query([q], evidence=ev, show_progress=False)
                    bp_q.append([case, q])
                    bp_vals.append(np.asarray(res.values, dtype=np.float64).reshape(-1))
        header = {"model": name, "ev_vars": ev_vars, "ve_queries": ve_q, "bp_queries": bp_q, "seed": 0, "reference": "pgmpy 1.0.0 (tristantreb/pgmpy), numpy backend, fp64"}
        path = os.path.join(OUT_DIR, f"ref_{name}.npz")
        np.savez_compressed(
            path,
            header=np.array(json.dumps(header)),
            ev_states=states,
            ve_values=np.concatenate(ve_vals) if ve_vals else np.zeros(0),
            ve_sizes=np.array([v.size for v in ve_vals], dtype=np.int32),
            bp_values=np.concatenate(bp_vals) if bp_vals else np.zeros(0),
            bp_sizes=np.array([v.size for v in bp_vals], dtype=np.int32),
        )
        print(f"{name}: {len(ve_q)} VE + {len(bp_q)} BP reference posteriors in {time.time() - t0:.1f}s -> {path}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(SPEC))
