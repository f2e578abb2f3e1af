"""Smallest exercise of k_contract_tc32 (pgx_tc32.cu: tcgen05 TF32x3, fp32 mode): one CPT-times-message step on synthetic
tables against numpy in fp64. Run under a short `timeout` whenever the kernel changes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pgmpy_b200.plan import PlanBuilder
from pgmpy_b200.engine import CompiledPlan

def case(M, N, K, Z, B, seed=0):
    rng = np.random.default_rng(seed)
    card = {"x": M, "y": N, "s": K, "z": Z, "e": 3}
    b = PlanBuilder(card, ["e"])
    q0 = b.add_const(["e", "s", "z", "y"], rng.random((3, K, Z, N)))
    pc = b.add_const(["z", "x", "s"], rng.random((Z, M, K)))
    q = b.contract([q0], ["s", "z", "y"])
    out = b.contract([pc, q], ["z", "x", "y"], optimize=False, split=False)
    b.emit(out, False)
    plan = b.finalize()
    cp = CompiledPlan(plan, dtype="float32")
    cp.set_mode("stepwise")
    ev = rng.integers(0, 3, size=(B, 1)).astype(np.int32)
    res = {}
    for tag, tc in (("ffma", False), ("tcgen05", True)):
        cp.set_tc32(tc)
        res[tag] = cp.run_host(ev).astype(np.float64)
        res[tag + "_n"] = cp.last_tc_steps
    blob = plan.const_blob
    Q0 = blob[q0.offset:q0.offset + q0.size].reshape(3, K, Z, N)
    PC = blob[pc.offset:pc.offset + pc.size].reshape(Z, M, K)
    want = np.stack([np.einsum("zxs,szy->zxy", PC, Q0[e]).reshape(-1) for e in ev[:, 0]])
    errs = {k: float(np.abs(v / want - 1).max()) for k, v in res.items() if not k.endswith("_n")}
    print(f"M{M} N{N} K{K} Z{Z} B={B}: tcgen05 steps {res['tcgen05_n']} (ffma run: {res['ffma_n']}), rel err", {k: f"{e:.1e}" for k, e in errs.items()}, flush=True)
    assert res["tcgen05_n"] == 1 and res["ffma_n"] == 0 and all(e <= 1e-5 for e in errs.values()), errs

if __name__ == "__main__":
    case(165, 85, 11, 1, 256)
    case(126, 55, 16, 1, 200)
    case(100, 39, 2, 1, 130)
    case(72, 20, 19, 3, 128)
    case(160, 7, 33, 2, 300)
    print("tc32_smoke ok")
