timeout 90 python tools/tc32_smoke.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "tcgen05" 2>&1 | grep -E "fp32 vs|passed|failed|Error|assert" | head
timeout 300 python - <<'PY'
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.inference import BeliefPropagation
for name, B in (("diabetes", 2048), ("munin", 256)):
    m = px.get_example_model(name)
    ev_vars, states = sample_evidence(m, B, 8, seed=1)
    bp = BeliefPropagation(m, dtype="float32")
    cp = bp.marginals_plan(ev_vars); cp.set_mode("stepwise")
    ev = torch.from_numpy(states).cuda()
    out = torch.empty((B, cp.out_elems), dtype=torch.float32, device="cuda")
    for tc in (False, True):
        cp.set_tc32(tc)
        for _ in range(4): cp.run(ev, out=out)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): cp.run(ev, out=out)
        b.record(); torch.cuda.synchronize()
        print(f"{name} B={B} float32 tc32={tc}: {a.elapsed_time(b)/5:.2f} ms  tc steps {cp.last_tc_steps} staged {cp.last_staged_steps}", flush=True)
PY
