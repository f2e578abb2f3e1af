"""Calibrated clique and sepset beliefs from the UNMODIFIED reference.

TEST INFRASTRUCTURE; build container only:  python -m oracle.make_golden_beliefs
alarm and hepar2: pgmpy.models.JunctionTree built from OUR min-fill cliques and clique potentials (the reference's own
triangulation yields 2e7..2e8-entry cliques on alarm), BeliefPropagation(jt).calibrate(), then
get_clique_beliefs() / get_sepset_beliefs() -> tests/golden/beliefs_<model>.npz (variables in OUR clique order).
The reference stops iterating when np.allclose accepts every sepset (ExactInference.py:807-895): its beliefs are exact
to about 1e-8 relative, which is the tolerance of the test that reads this file.
"""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.factors.discrete import DiscreteFactor as RefDF
    from pgmpy.inference import BeliefPropagation as RefBP
    from pgmpy.models import JunctionTree as RefJT

    import pgmpy_b200 as px
    from pgmpy_b200.planner import JTStructure

    for name in ("alarm", "hepar2"):
        m = px.get_example_model(name)
        jt = JTStructure.from_model(m)
        rjt = RefJT()
        for c in jt.cliques:
            rjt.add_node(c)
        for a, b in jt.edges:
            rjt.add_edge(jt.cliques[a], jt.cliques[b])
        for c, p in zip(jt.cliques, jt.potentials):
            rjt.add_factors(RefDF(list(c), [jt.card[v] for v in c], p, state_names={v: m.states[v] for v in c}))
        rbp = RefBP(rjt)
        rbp.calibrate()
        cb, sb = rbp.get_clique_beliefs(), rbp.get_sepset_beliefs()
        arrays, cliques, sepsets = {}, [], []
        for i, c in enumerate(jt.cliques):
            f = cb[c]
            perm = [f.variables.index(v) for v in c]
            arrays[f"c{i}"] = np.ascontiguousarray(np.transpose(np.asarray(f.values, dtype=np.float64), perm))
            cliques.append(list(c))
        for k, (a, b) in enumerate(jt.edges):
            f = sb[frozenset((jt.cliques[a], jt.cliques[b]))]
            vars_ = sorted(f.variables, key=str)
            perm = [f.variables.index(v) for v in vars_]
            arrays[f"s{k}"] = np.ascontiguousarray(np.transpose(np.asarray(f.values, dtype=np.float64), perm))
            sepsets.append({"a": list(jt.cliques[a]), "b": list(jt.cliques[b]), "vars": vars_})
        path = os.path.join(OUT_DIR, f"beliefs_{name}.npz")
        header = {"model": name, "cliques": cliques, "sepsets": sepsets,
                  "reference": "pgmpy 1.0.0 BeliefPropagation(JunctionTree of our cliques/potentials).calibrate()"}
        np.savez_compressed(path, header=np.array(json.dumps(header)), **arrays)
        print(name, len(cliques), "cliques", len(sepsets), "sepsets ->", path)


if __name__ == "__main__":
    main()
