"""Per-step device timing of a stepwise pass (CUDA events, pgx_profile_steps). Usage:
    python tools/profile_steps.py munin 64 [generic]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.inference import BeliefPropagation

name, B = sys.argv[1], int(sys.argv[2])
kern = sys.argv[3] if len(sys.argv) > 3 else "auto"
m = px.get_example_model(name)
bp = BeliefPropagation(m)
ev_vars, states = sample_evidence(m, B, 5 if name == "alarm" else 8, seed=1)
cp = bp.marginals_plan(ev_vars)
cp.set_mode("stepwise", 0, "auto", kern)
ev = torch.from_numpy(states).cuda()
cp.profile_steps(ev)
rows = cp.profile_steps(ev)
tot = sum(r[0] for r in rows)
print(f"{name} B={B} kernel={kern}: {len(rows)} steps, {tot:.2f} ms total (sum of per-step event times)")
bins = [(0, 0.004), (0.004, 0.008), (0.008, 0.02), (0.02, 0.1), (0.1, 1), (1, 1e9)]
for lo, hi in bins:
    sel = [r for r in rows if lo <= r[0] < hi]
    print(f"  steps with {lo*1e3:.0f}-{hi*1e3:.0f} us: {len(sel):5d}  sum {sum(r[0] for r in sel):8.2f} ms  alg GB {sum(r[4] for r in sel)/1e9:8.2f}")
print("  top steps: ms, out, sum, K, alg GB/s, level")
for r in sorted(rows, reverse=True)[:18]:
    print(f"   {r[0]:8.3f} {r[1]:9d} {r[2]:7d} {r[3]:2d} {r[4]/r[0]/1e6:9.0f} {r[5]:3d}")
