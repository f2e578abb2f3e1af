"""Whole-pass time of the HBM-resident configs with the production path (CUDA-graph replay), CUDA events."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.inference import BeliefPropagation

peak = 6544.7
for name, B in (("diabetes", 2048), ("munin", 256), ("pathfinder", 16384)):
    if len(sys.argv) > 1 and name not in sys.argv[1:]:
        continue
    m = px.get_example_model(name)
    ev_vars, states = sample_evidence(m, B, 8, seed=1)
    bp = BeliefPropagation(m)
    cp = bp.marginals_plan(ev_vars)
    ev = torch.from_numpy(states).cuda()
    out = torch.empty((B, cp.out_elems), dtype=torch.float64, device="cuda")
    for _ in range(4):
        cp.run(ev, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        cp.run(ev, out=out)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 5
    alg = cp.plan.algorithmic_bytes(B)
    print(f"{name} B={B}: {ms:.3f} ms  {alg/ms/1e6:.0f} GB/s  frac {alg/ms/1e6/peak:.3f}  graph={cp.last_graph} launches={cp.last_launches}", flush=True)
