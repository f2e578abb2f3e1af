"""Step census of a compiled junction-tree plan (host only, no GPU): where the algorithmic bytes and the
multiply-adds of a plan go, by step class. Used to decide which kernel variant a step should get.

    python tools/plan_census.py munin [k] [top]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan
from pgmpy_b200.plan import KIND_WORK


def census(name, k=8, top=25):
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    ev_vars, _ = sample_evidence(m, 1, k, seed=1)
    plan = compile_jt_plan(jt, ev_vars)
    card = plan.card
    rows = []
    for st in plan.steps:
        out = st.out.size
        ssz = 1
        for v in st.sum_vars:
            ssz *= card[v]
        work = [t.size for t, d in st.operands if t.kind == KIND_WORK]
        const = [t.size for t, d in st.operands if t.kind != KIND_WORK]
        ev_const = sum(1 for t, d in st.operands if t.kind != KIND_WORK and any(v in plan.ev_vars for v in t.vars))
        outset = set(st.out.vars)
        # does a work operand depend on output variables / on summed variables?
        dep = []
        for t, d in st.operands:
            fv = [v for v in t.vars if v not in plan.ev_vars]
            o_part = int(np.prod([card[v] for v in fv if v in outset], dtype=np.int64))
            s_part = int(np.prod([card[v] for v in fv if v not in outset], dtype=np.int64))
            dep.append(("W" if t.kind == KIND_WORK else "C") + ("d" if d else "") + f"{o_part}x{s_part}")
        bytes_ = 8 * (out + sum(work))
        rows.append(dict(out=out, sum=ssz, K=len(st.operands), work=work, const=const, ev_const=ev_const, bytes=bytes_,
                         fma=out * ssz * len(st.operands), dep=" ".join(dep), level=st.level, div=any(d for _, d in st.operands)))
    tot_b = sum(r["bytes"] for r in rows)
    tot_f = sum(r["fma"] for r in rows)
    print(f"{name}: {len(rows)} steps, levels {1 + max(r['level'] for r in rows)}, ws_entries {plan.ws_entries}, "
          f"alg bytes/set {plan.algorithmic_bytes(1)/1e6:.2f} MB (work part {tot_b/1e6:.2f} MB), operand loads/set {tot_f/1e6:.2f} M, "
          f"const blob {plan.const_blob.size*8/1e6:.1f} MB")
    # classes
    def cls(r):
        if r["sum"] == 1:
            return "product (S=1)" + (" div" if r["div"] else "")
        if len(r["work"]) == 1 and not r["const"]:
            return "marginalise one work table"
        if not r["work"]:
            return "const only"
        return f"contract {len(r['work'])}W+{len(r['const'])}C"
    agg = {}
    for r in rows:
        c = cls(r)
        a = agg.setdefault(c, [0, 0, 0])
        a[0] += 1
        a[1] += r["bytes"]
        a[2] += r["fma"]
    for c, (n, b, f) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"  {c:32s} steps {n:5d}  bytes {b/1e6:9.2f} MB ({100*b/tot_b:5.1f} %)  loads {f/1e6:9.2f} M ({100*f/tot_f:5.1f} %)")
    print("  top steps by bytes: out, sum, K, bytes MB, loads M, level, operands (kind o-part x s-part)")
    for r in sorted(rows, key=lambda r: -r["bytes"])[:top]:
        print(f"    {r['out']:9d} {r['sum']:6d} {r['K']:2d} {r['bytes']/1e6:8.2f} {r['fma']/1e6:8.2f} {r['level']:3d}  {r['dep']}")
    print("  top steps by loads:")
    for r in sorted(rows, key=lambda r: -r["fma"])[:top]:
        print(f"    {r['out']:9d} {r['sum']:6d} {r['K']:2d} {r['bytes']/1e6:8.2f} {r['fma']/1e6:8.2f} {r['level']:3d}  {r['dep']}")
    return plan


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "munin"
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    census(name, k, top)
