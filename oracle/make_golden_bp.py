"""BP-mode golden posteriors from the UNMODIFIED reference, exact to fp64 rounding (no calibration loop).

TEST INFRASTRUCTURE; build container only:   python -m oracle.make_golden_bp [model ...] [--jobs N]

BP mode is `normalise( sum_{all \\ (q u E)} prod_{all v} CPT_v[E = e] )`: every factor, no pruning.
pgmpy's own `BeliefPropagation.query` reaches that value only up to its iterate-until-allclose calibration
(pgmpy/inference/ExactInference.py:807-895; 2e-8 off on hepar2), so it cannot pin a 1e-12 comparison. The reference
computes the SAME closed form exactly through its in-tree classic variable elimination on a non-Bayesian model:

    mm = model.to_markov_model()                  # pgmpy/models/DiscreteBayesianNetwork.py:510-537: moralise, every
                                                  # CPD becomes a factor via to_factor(), nothing dropped
    VariableElimination(mm).query([q], evidence, elimination_order=<explicit list>)

For a DiscreteMarkovNetwork `query` skips `_prune_bayesian_model` (ExactInference.py:339-345), validates the explicit
list (:91-118), runs the factor_product / marginalize loop (:200-215) and returns the UN-normalised product of what is
left (:225-229); we normalise it (values / values.sum(), DiscreteFactor.py:530). The elimination order is ours
(min-fill over all factors, query variable last): pgmpy's own heuristics exhaust memory on diabetes / munin
 and order changes results only at ~1e-16.

Writes tests/golden/refbp_<model>.npz: evidence variables / states (forward-sampled, seed 0), query list, posteriors.
Protocol: 256 evidence sets for alarm / hepar2 / win95pts, 64 for pathfinder, 16 for diabetes / munin.
"""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "tests", "golden")

# model -> (k evidence vars, evidence sets, query variables per set (None = every unobserved variable))
SPEC = {
    "asia": (2, 32, None),
    "cancer": (1, 16, None),
    "sachs": (3, 32, None),
    "child": (4, 64, None),
    "alarm": (5, 256, None),
    "hepar2": (8, 256, 8),
    "win95pts": (8, 256, 8),
    "pathfinder": (8, 64, 8),
    "munin": (8, 16, 8),
    "diabetes": (8, 16, 8),
}

_W = {}


def _init(name):
    from oracle.ref_loader import load_reference

    load_reference()
    from pgmpy.inference import VariableElimination as RefVE
    from pgmpy.utils import get_example_model as ref_model

    import pgmpy_b200 as px

    rm = ref_model(name)
    _W["ve"] = RefVE(rm.to_markov_model())
    _W["ve_bn"] = RefVE(rm)  # VE mode: the Bayesian network itself, with the reference's own pruning
    _W["m"] = px.get_example_model(name)


def _one(task):
    from pgmpy_b200.evidence import states_to_names
    from pgmpy_b200.planner import compile_ve_plan

    case, q, ev_vars, row, mode = task
    m = _W["m"]
    ev = states_to_names(m, ev_vars, row)
    with np.errstate(all="ignore"):
        if mode == "ve":
            # VE mode (pruned, normalised by the reference itself): the SURVEY 8c oracle protocol
            order = list(compile_ve_plan(m, [q], ev_vars).meta["order"])
            res = _W["ve_bn"].query([q], evidence=ev, elimination_order=order, show_progress=False)
            vals = np.asarray(res.values, dtype=np.float64).reshape(-1)
        else:
            order = list(compile_ve_plan(m, [q], ev_vars, prune=False).meta["order"])
            res = _W["ve"].query([q], evidence=ev, elimination_order=order, show_progress=False)
            vals = np.asarray(res.values, dtype=np.float64).reshape(-1)
            vals = vals / vals.sum()
    return case, q, vals


def main(argv):
    jobs = 8
    mode = "bp"
    names = []
    it = iter(argv)
    for a in it:
        if a == "--jobs":
            jobs = int(next(it))
        elif a == "--mode":
            mode = next(it)  # "bp" (default) or "ve": extra VE-mode goldens (refve_*) at the same protocol sizes
        else:
            names.append(a)
    import multiprocessing as mp

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence

    for name in names or list(SPEC):
        k, n_cases, n_q = SPEC[name]
        t0 = time.time()
        m = px.get_example_model(name)
        ev_vars, states = sample_evidence(m, n_cases, k, seed=0)
        free = [v for v in sorted(m.nodes(), key=str) if v not in ev_vars]
        rng = np.random.default_rng(2718)
        tasks = []
        for case in range(n_cases):
            qs = free if n_q is None else [free[i] for i in sorted(rng.choice(len(free), n_q, replace=False))]
            tasks += [(case, q, ev_vars, states[case], mode) for q in qs]
        with mp.get_context("fork").Pool(min(jobs, len(tasks)), initializer=_init, initargs=(name,)) as pool:
            res = pool.map(_one, tasks, chunksize=max(1, len(tasks) // (jobs * 8)))
        header = {"model": name, "ev_vars": ev_vars, "queries": [[c, q] for c, q, _ in res], "seed": 0,
                  "reference": "pgmpy 1.0.0 (tristantreb/pgmpy), numpy backend, fp64: " + (
                      "VariableElimination(model).query([q], evidence, elimination_order=<explicit min-fill list>) incl. pruning"
                      if mode == "ve" else
                      "VariableElimination(model.to_markov_model()).query([q], evidence, elimination_order=<explicit min-fill list>), normalised")}
        path = os.path.join(OUT_DIR, f"ref{mode}_{name}.npz")
        np.savez_compressed(path, header=np.array(json.dumps(header)), ev_states=states,
                            values=np.concatenate([v for _, _, v in res]),
                            sizes=np.array([v.size for _, _, v in res], dtype=np.int32))
        print(f"{name}: {len(res)} {mode.upper()}-mode reference posteriors ({n_cases} evidence sets) in {time.time() - t0:.1f}s -> {path}", flush=True)


if __name__ == "__main__":
    main(sys.argv[1:])
