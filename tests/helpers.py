"""Shared fixtures: the small networks and known answers of the reference's own tests, golden loaders."""
import ctypes as C
import json
import os

import numpy as np

import pgmpy_b200 as px
from pgmpy_b200 import DiscreteBayesianNetwork, TabularCPD

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def six_node_net():
    """pgmpy/tests/test_inference/test_ExactInference.py:24-49 (values checked with SAMIAM there)."""
    m = DiscreteBayesianNetwork([("A", "J"), ("R", "J"), ("J", "Q"), ("J", "L"), ("G", "L")])
    m.add_cpds(
        TabularCPD("A", 2, [[0.2], [0.8]]),
        TabularCPD("G", 2, [[0.6], [0.4]]),
        TabularCPD("J", 2, [[0.9, 0.6, 0.7, 0.1], [0.1, 0.4, 0.3, 0.9]], ["A", "R"], [2, 2]),
        TabularCPD("L", 2, [[0.9, 0.45, 0.8, 0.1], [0.1, 0.55, 0.2, 0.9]], ["J", "G"], [2, 2]),
        TabularCPD("Q", 2, [[0.9, 0.2], [0.1, 0.8]], ["J"], [2]),
        TabularCPD("R", 2, [[0.4], [0.6]]),
    )
    return m


# (variables, evidence, expected joint in the order of `variables`), test_ExactInference.py:63-141
SIX_NODE_ANSWERS = [
    (["J"], {}, [0.416, 0.584]),
    (["J", "Q"], {}, [[0.3744, 0.0416], [0.1168, 0.4672]]),
    (["J"], {"A": 0, "R": 1}, [0.6, 0.4]),
    (["J", "Q"], {"A": 0, "R": 0, "G": 0, "L": 1}, [[0.73636364, 0.08181818], [0.03636364, 0.14545455]]),
]


def snow_net():
    """test_ExactInference.py:385-427 — named states."""
    m = DiscreteBayesianNetwork([("Snow", "Risk"), ("Snow", "Traffic"), ("Traffic", "Late"), ("Risk", "Late")])
    m.add_cpds(
        TabularCPD("Snow", 2, [[0.4], [0.6]], state_names={"Snow": ["yes", "no"]}),
        TabularCPD("Risk", 2, [[0.8, 0.4], [0.2, 0.6]], ["Snow"], [2], state_names={"Snow": ["yes", "no"], "Risk": ["yes", "no"]}),
        TabularCPD("Traffic", 2, [[0.4, 0.65], [0.6, 0.35]], ["Snow"], [2], state_names={"Traffic": ["normal", "slow"], "Snow": ["yes", "no"]}),
        TabularCPD(
            "Late", 2, [[0.45, 0.85, 0.1, 0.7], [0.55, 0.15, 0.90, 0.30]], ["Risk", "Traffic"], [2, 2],
            state_names={"Late": ["yes", "no"], "Traffic": ["normal", "slow"], "Risk": ["yes", "no"]},
        ),
    )
    return m


# test_ExactInference.py:429-447
SNOW_ANSWERS = [
    (["Snow"], {"Traffic": "slow"}, [0.533333, 0.466667]),
    (["Risk"], {"Traffic": "slow"}, [0.613333, 0.386667]),
    (["Late"], {"Traffic": "slow"}, [0.7920, 0.2080]),
]
# virtual evidence on Traffic = [0.3, 0.7] (:529-560) and additionally Risk = [0.7, 0.3] (:571-629)
SNOW_VIRTUAL_1 = [(["Snow"], [0.45, 0.55]), (["Risk"], [0.58, 0.42]), (["Late"], [0.61625, 0.38375]), (["Traffic"], [0.34375, 0.65625])]
SNOW_VIRTUAL_2 = [
    (["Snow"], [0.52443609, 0.47556391]),
    (["Risk"], [0.76315789, 0.23684211]),
    (["Traffic"], [0.32730263, 0.67269737]),
    (["Late"], [0.66480263, 0.33519737]),
]


def load_golden(name):
    """tests/golden/ref_<name>.npz -> dict(ev_vars, ev_states, ve=[(case, q, values)], bp=[...])."""
    path = os.path.join(GOLDEN, f"ref_{name}.npz")
    with np.load(path) as z:
        hdr = json.loads(str(z["header"]))
        out = {"ev_vars": hdr["ev_vars"], "ev_states": z["ev_states"].astype(np.int32)}
        for key in ("ve", "bp"):
            vals, sizes = z[f"{key}_values"], z[f"{key}_sizes"]
            items, off = [], 0
            for (case, q), n in zip(hdr[f"{key}_queries"], sizes):
                items.append((int(case), q, vals[off : off + n]))
                off += int(n)
            out[key] = items
    return out


def golden_models():
    return sorted(f[4:-4] for f in os.listdir(GOLDEN) if f.startswith("ref_") and f.endswith(".npz"))


def rel_err(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    denom = np.maximum(np.abs(want), 1e-300)
    err = np.abs(got - want) / denom
    err = np.where((want == 0) & (got == 0), 0.0, err)
    return float(np.max(err)) if err.size else 0.0


_hostsim = None


def hostsim_run(plan, ev, dtype=np.float64, use_run=True):
    """Runs the device element function (pgx_step.cuh) compiled for the host over a packed plan."""
    global _hostsim
    if _hostsim is None:
        _hostsim = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "_hostsim.so"))
    ev = np.ascontiguousarray(np.asarray(ev, dtype=np.int32)).reshape(-1, len(plan.ev_vars))
    B = ev.shape[0]
    ldb = (B + 31) // 32 * 32 if B >= 32 else 1 << max(0, (B - 1).bit_length())
    pool = np.ascontiguousarray(plan.pool)
    cst = np.ascontiguousarray(plan.const_blob.astype(dtype))
    ws = np.zeros(plan.ws_entries * ldb, dtype=dtype)
    out = np.zeros((B, plan.out_elems), dtype=dtype)
    fn = _hostsim.hostsim_run_f64 if dtype == np.float64 else _hostsim.hostsim_run_f32
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    fn(p(pool), p(cst), p(ev), p(ws), p(out), C.c_int64(B), C.c_int64(ldb), C.c_int(1 if use_run else 0))
    return out


# The reference's BeliefPropagation.query stops calibrating when DiscreteFactor.__eq__ (np.allclose, rtol 1e-5 /
# atol 1e-8, pgmpy/inference/ExactInference.py:807-895) accepts every sepset, so ITS output is only that exact
# (measured: 2e-15 on alarm, 2e-8 on hepar2). Parity is therefore pinned at 1e-12 against the refbp_* goldens (the
# reference's exact classic VE over all factors, oracle/make_golden_bp.py); the BeliefPropagation.query goldens are a
# secondary check held to the reference's own stopping rule.
BP_QUERY_REFERENCE_RESIDUAL = 1e-6


def load_golden_bp(name, kind="bp"):
    """tests/golden/ref{bp,ve}_<name>.npz -> dict(ev_vars, ev_states, items=[(case, q, values)]) — BP-mode (or, kind="ve",
    additional VE-mode) posteriors of the unmodified reference at the SURVEY 8d protocol sizes, exact to rounding
    (oracle/make_golden_bp.py)."""
    path = os.path.join(GOLDEN, f"ref{kind}_{name}.npz")
    with np.load(path) as z:
        hdr = json.loads(str(z["header"]))
        vals, sizes = z["values"], z["sizes"]
        items, off = [], 0
        for (case, q), n in zip(hdr["queries"], sizes):
            items.append((int(case), q, vals[off : off + n]))
            off += int(n)
        return {"ev_vars": hdr["ev_vars"], "ev_states": z["ev_states"].astype(np.int32), "items": items}


def golden_bp_models(kind="bp"):
    pre = f"ref{kind}_"
    return sorted(f[len(pre):-4] for f in os.listdir(GOLDEN) if f.startswith(pre) and f.endswith(".npz"))


def hostsim_micro_run(plan, ev):
    """Walks the table-driven microprogram (pgx_fused.cuh::build_micro) on the CPU. Returns (out, n_levels)."""
    global _hostsim
    if _hostsim is None:
        _hostsim = C.CDLL(os.path.join(ROOT, "tests", "hostsim", "_hostsim.so"))
    ev = np.ascontiguousarray(np.asarray(ev, dtype=np.int32)).reshape(-1, max(1, len(plan.ev_vars)))
    B = ev.shape[0]
    ldb = (B + 31) // 32 * 32 if B >= 32 else 1 << max(0, (B - 1).bit_length())
    pool = np.ascontiguousarray(plan.pool)
    cst = np.ascontiguousarray(plan.const_blob)
    ws = np.zeros(plan.ws_entries * ldb)
    out = np.zeros((B, plan.out_elems))
    nl = C.c_int32(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = _hostsim.hostsim_micro_f64(p(pool), p(cst), p(ev), p(ws), p(out), C.c_int64(B), C.c_int64(ldb), C.byref(nl))
    if rc != 0:
        raise RuntimeError("plan too large for a microprogram")
    return out, nl.value
