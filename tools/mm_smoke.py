"""Smallest possible exercise of k_contract_mm (pgx_mm.cu): one matrix-product-shaped step on synthetic tables, against
numpy. Run under a short `timeout` first whenever the kernel's pipeline changes: a deadlock must not cost GPU minutes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pgmpy_b200.plan import PlanBuilder
from pgmpy_b200.engine import CompiledPlan

def case(M, N, K, Z, const_p, B, dtype="float64", seed=0):
    rng = np.random.default_rng(seed)
    card = {"x": M, "y": N, "s": K, "z": Z, "e": 3}
    b = PlanBuilder(card, ["e"])
    # batch-dependent operands: evidence variable e selects a slice of a constant table
    p0 = b.add_const(["e", "z", "x", "s"], rng.random((3, Z, M, K)))
    q0 = b.add_const(["e", "s", "z", "y"], rng.random((3, K, Z, N)))
    pc = b.add_const(["z", "x", "s"], rng.random((Z, M, K)))
    p = pc if const_p else b.contract([p0], ["z", "x", "s"])
    q = b.contract([q0], ["s", "z", "y"])
    out = b.contract([p, q], ["z", "x", "y"], optimize=False, split=False)
    b.emit(out, False)
    plan = b.finalize()
    cp = CompiledPlan(plan, dtype=dtype)
    cp.set_mode("stepwise")
    ev = rng.integers(0, 3, size=(B, 1)).astype(np.int32)
    res = {}
    for tag, stage, mma in (("stream", 0, 1), ("mm+mma", 1, 1), ("mm", 1, 0)):
        cp.set_stage(stage); cp.set_mma(bool(mma))
        res[tag] = cp.run_host(ev).astype(np.float64)
        res[tag + "_n"] = cp.last_staged_steps
    blob = plan.const_blob
    P0 = blob[p0.offset:p0.offset + p0.size].reshape(3, Z, M, K); Q0 = blob[q0.offset:q0.offset + q0.size].reshape(3, K, Z, N)
    PC = blob[pc.offset:pc.offset + pc.size].reshape(Z, M, K)
    want = np.stack([np.einsum("zxs,szy->zxy", PC if const_p else P0[e], Q0[e]).reshape(-1) for e in ev[:, 0]])
    tol = 1e-13 if dtype == "float64" else 1e-5
    errs = {k: float(np.abs(v / want - 1).max()) for k, v in res.items() if not k.endswith("_n")}
    print(f"M{M} N{N} K{K} Z{Z} const_p={const_p} B={B} {dtype}: staged steps {res['mm_n']}, rel err", {k: f"{e:.1e}" for k, e in errs.items()}, flush=True)
    assert res["mm_n"] >= 1 and all(e <= tol for e in errs.values()), errs

if __name__ == "__main__":
    case(21, 55, 96, 1, False, 70)
    case(21, 55, 96, 1, True, 70)
    case(16, 49, 7, 50, False, 40)
    case(165, 85, 11, 1, True, 64)
    case(5, 187, 10, 1, False, 33)
    case(13, 10, 17, 11, False, 64, "float32")
    case(100, 39, 2, 1, True, 64, "float32")
    print("mm_smoke ok")
