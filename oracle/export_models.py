"""Export the reference's bundled bnlearn example models to compact fixtures.

TEST INFRASTRUCTURE; run in the build container only (needs /root/reference):

    python -m oracle.export_models [names...]

Each model is loaded with the reference's own reader (`pgmpy.utils.get_example_model`,
/root/reference/pgmpy/utils/utils.py:16 -> readwrite/BIF.py:361) and written to
`pgmpy_b200/data/models/<name>.npz` as: a JSON header (node order, edges, per-CPD variable order =
child first then parents in BIF order, cardinalities, state names) + one packed fp64 value blob in
the reference's C-order layout (DiscreteFactor.py:91-127). No reference source is copied.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "pgmpy_b200", "data", "models")
DEFAULT = ["asia", "cancer", "sachs", "child", "alarm", "hepar2", "win95pts", "pathfinder", "munin", "diabetes"]


def export(name):
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.utils import get_example_model

    model = get_example_model(name)
    nodes = list(model.nodes())
    header = {"name": name, "nodes": nodes, "edges": [list(e) for e in model.edges()], "cpds": []}
    blobs = []
    off = 0
    for node in nodes:
        cpd = model.get_cpds(node)
        vals = np.ascontiguousarray(np.asarray(cpd.values, dtype=np.float64)).reshape(-1)
        header["cpds"].append(
            {
                "variable": cpd.variable,
                "variables": list(cpd.variables),
                "cardinality": [int(c) for c in cpd.cardinality],
                "state_names": {v: list(cpd.state_names[v]) for v in cpd.variables},
                "offset": off,
                "size": int(vals.size),
            }
        )
        blobs.append(vals)
        off += vals.size
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    np.savez_compressed(path, header=np.array(json.dumps(header)), values=np.concatenate(blobs))
    return path, len(nodes), off


if __name__ == "__main__":
    for n in sys.argv[1:] or DEFAULT:
        p, nn, sz = export(n)
        print(f"{n}: {nn} nodes, {sz} CPT entries -> {p} ({os.path.getsize(p)} bytes)")
