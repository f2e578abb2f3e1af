"""Per-launch device timing of a stepwise pass with the production schedule (pgx_profile_launches), next to the
algorithmic bytes of the steps each launch serves. With `ncu` set, runs plain launches (no CUDA graph, no events) so
that an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list of this command
lines up one to one with the table printed by the plain run.

    python tools/launch_profile.py munin 256 [top] [ncu]
"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.inference import BeliefPropagation

name, B = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ncu = "ncu" in sys.argv[4:]
m = px.get_example_model(name)
bp = BeliefPropagation(m, dtype="float32" if "f32" in sys.argv[4:] else "float64")
ev_vars, states = sample_evidence(m, B, 5 if name == "alarm" else 8, seed=1)
cp = bp.marginals_plan(ev_vars)
cp.set_mode("stepwise")
if "nostage" in sys.argv[4:]:
    cp.set_stage(False)
if "nomma" in sys.argv[4:]:
    cp.set_mma(False)
print("plan:", cp.plan.meta.get("distribute"), "factorized" if cp.plan.meta.get("factorized") else "dense", cp.plan.n_steps, "steps")
ev = torch.from_numpy(states).cuda()
if ncu:
    cp.set_graph(False)
    cp.run(ev)
    torch.cuda.synchronize()
    cp.run(ev)
    torch.cuda.synchronize()
    print("launches per pass:", cp.last_launches)
    sys.exit(0)
cp.profile_launches(ev)
rows = cp.profile_launches(ev)
tot = sum(r["ms"] for r in rows)
alg = sum(r["alg_bytes"] for r in rows)
print(f"{name} B={B}: {len(rows)} launches, {tot:.3f} ms (sum of per-launch event times), alg {alg/1e9:.2f} GB -> {alg/tot/1e6:.0f} GB/s")
card = cp.plan.card
def describe(si):
    st = cp.plan.steps[si]
    ssz = 1
    for v in st.sum_vars:
        ssz *= card[v]
    return f"{st.out.size}x{ssz}:" + "".join(("W" if t.kind == 1 else "C") + ("d" if d else "") for t, d in st.operands)
print("  launch  ms  steps  alg GB  GB/s  biggest steps (out x sum : operands)")
for i, r in enumerate(rows):
    big = sorted(r["steps"], key=lambda si: -cp.plan.steps[si].out.size * max(1, len(cp.plan.steps[si].operands)))[:4]
    r["desc"] = " ".join(describe(si) for si in big)
    r["i"] = i
for r in sorted(rows, key=lambda r: -r["ms"])[:top]:
    print(f"  {r['i']:4d} {r['ms']:8.3f} {len(r['steps']):5d} {r['alg_bytes']/1e9:8.3f} {r['alg_bytes']/max(r['ms'],1e-6)/1e6:7.0f}  {r['desc']}")
os.makedirs("gpurun_out", exist_ok=True)
json.dump([{k: r[k] for k in ("i", "ms", "alg_bytes", "desc")} | {"n_steps": len(r["steps"])} for r in rows],
          open(f"gpurun_out/launch_profile_{name}_{B}{'_nostage' if 'nostage' in sys.argv[4:] else ''}{'_nomma' if 'nomma' in sys.argv[4:] else ''}.json", "w"))
