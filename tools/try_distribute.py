"""Compare distribute variants of the junction-tree plan on the GPU: python tools/try_distribute.py diabetes 2048"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgmpy_b200 as px
from pgmpy_b200.engine import CompiledPlan
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan

name, B = sys.argv[1], int(sys.argv[2])
m = px.get_example_model(name)
jt = JTStructure.from_model(m)
ev_vars, states = sample_evidence(m, B, 8, seed=1)
ev = torch.from_numpy(states).cuda()
for d in sys.argv[3:] or ["ss", "belief", "divide"]:
    gemm = d.endswith("-gemm")
    vec2 = not d.endswith("-novec")
    plan = compile_jt_plan(jt, ev_vars, distribute=d.replace("-gemm", "").replace("-novec", ""))
    cp = CompiledPlan(plan)
    cp.set_gemm_tile(gemm)
    cp.set_vec2(vec2)
    out = torch.empty((B, cp.out_elems), dtype=torch.float64, device="cuda")
    for _ in range(2):
        cp.run(ev, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        cp.run(ev, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(json.dumps({"model": name, "B": B, "distribute": d, "steps": plan.n_steps, "loads": plan.operand_loads(),
                      "alg_MB_per_set": plan.algorithmic_bytes(B) / B / 1e6, "ms": ms, "q_per_s": B / ms * 1e3,
                      "mode": cp.last_mode, "launches": cp.last_launches}), flush=True)
