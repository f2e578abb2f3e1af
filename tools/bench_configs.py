"""Runs the BASELINE.json configurations other than the headline one and prints one JSON line each
(device-resident throughput, CUDA events; parity of these paths is asserted in tests/test_gpu_parity.py).

    python tools/bench_configs.py [alarm_ve] [mixed_ve] [large] [munin]
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pgmpy_b200 as px
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.inference import BeliefPropagation, VariableElimination

PEAK = 6544.7
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, warm=2, reps=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def alarm_ve(batch=1024):
    """configs[0]: alarm, VariableElimination.query all-variable marginals over 1024 evidence sets, fp64."""
    m = px.get_example_model("alarm")
    ve = VariableElimination(m)
    ev_vars, states = sample_evidence(m, batch, 5, seed=1)
    ev = torch.from_numpy(states).cuda()
    free = [v for v in m.nodes() if v not in ev_vars]
    t0 = time.perf_counter()
    plans = [ve._plan([q], ev_vars, True, None) for q in free]
    compile_s = time.perf_counter() - t0
    outs = [torch.empty((batch, cp.out_elems), dtype=torch.float64, device="cuda") for cp in plans]

    def run():
        for cp, o in zip(plans, outs):
            cp.run(ev, out=o)

    t0 = time.perf_counter()
    multi = ve.marginals_plan(ev_vars)
    multi_compile_s = time.perf_counter() - t0
    mout = torch.empty((batch, multi.out_elems), dtype=torch.float64, device="cuda")
    ms_multi = timed(lambda: multi.run(ev, out=mout))
    for cp, o, seg in zip(plans, outs, multi.plan.segments):
        cp.run(ev, out=o)
        assert torch.allclose(o, mout[:, seg.out_offset : seg.out_offset + seg.table.size], rtol=1e-13, atol=0)
    print(json.dumps({"config": "alarm VE all-variable marginals, ONE plan holding all 32 pruned queries", "batch": batch,
                      "steps": multi.plan.n_steps, "plan_compile_s": round(multi_compile_s, 2), "mode": multi.last_mode,
                      "variant": multi.last_variant, "ms_per_batch": ms_multi, "evidence_queries_per_sec": batch / ms_multi * 1e3,
                      "single_variable_queries_per_sec": batch * len(plans) / ms_multi * 1e3,
                      "alg_GBps": multi.plan.algorithmic_bytes(batch) / ms_multi / 1e6,
                      "frac_of_hbm_peak": multi.plan.algorithmic_bytes(batch) / ms_multi / 1e6 / PEAK}), flush=True)
    # the same plan specialised (one warp per 32 evidence sets runs all 32 pruned queries out of registers)
    try:
        ref_out = mout.clone()
        t0 = time.perf_counter()
        info = multi.specialize()
        spec_s = time.perf_counter() - t0
        ms_spec = timed(lambda: multi.run(ev, out=mout), warm=3, reps=20)
        assert torch.allclose(mout, ref_out, rtol=1e-12, atol=0)
        for big in (131072,):
            _, st_big = sample_evidence(m, big, 5, seed=2, evidence_vars=ev_vars)
            ev_big = torch.from_numpy(st_big).cuda()
            out_big = torch.empty((big, multi.out_elems), dtype=torch.float64, device="cuda")
            ms_big = timed(lambda: multi.run(ev_big, out=out_big), warm=3, reps=20)
        print(json.dumps({"config": "alarm VE all-variable marginals, ONE plan, specialised kernel", "batch": batch, "variant": multi.last_variant,
                          "ms_per_batch": ms_spec, "evidence_queries_per_sec": batch / ms_spec * 1e3, "specialize_s": round(spec_s, 2),
                          "ms_per_131072_sets": ms_big, "evidence_queries_per_sec_at_131072": big / ms_big * 1e3, "spec": info}), flush=True)
    except Exception as exc:
        print(json.dumps({"config": "alarm VE all-variable marginals, specialised", "error": str(exc)[:300]}), flush=True)
    ms = timed(run)
    alg = sum(cp.plan.algorithmic_bytes(batch) for cp in plans)
    print(json.dumps({"config": "alarm VE all-variable marginals (one pruned plan per query variable)", "batch": batch,
                      "plans": len(plans), "plan_compile_s": round(compile_s, 2), "ms_per_batch": ms,
                      "evidence_queries_per_sec": batch / ms * 1e3, "single_variable_queries_per_sec": batch * len(plans) / ms * 1e3,
                      "alg_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / PEAK}), flush=True)


def single_query_latency(name="alarm", n=200):
    """Drop-in API latency: VariableElimination.query / BeliefPropagation.query for ONE evidence set (plan cached),
    wall clock per call including H2D of the evidence, the launch sequence, D2H and DiscreteFactor construction.
    The reference needs ~5.4 ms per alarm VE query on a CPU core (BASELINE.md probe)."""
    m = px.get_example_model(name)
    ve, bp = VariableElimination(m), BeliefPropagation(m)
    ev_vars, states = sample_evidence(m, n, 5 if name == "alarm" else 8, seed=3)
    free = [v for v in m.nodes() if v not in ev_vars]
    rows = [{v: m.states[v][int(s)] for v, s in zip(ev_vars, st)} for st in states]
    res = {}
    for label, infer in (("ve", ve), ("bp", bp)):
        infer.query([free[0]], evidence=rows[0])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(n):
            infer.query([free[0]], evidence=rows[i])
        res[label] = (time.perf_counter() - t0) / n * 1e6
        # a new query variable = a new plan: compile + upload + first run
        t0 = time.perf_counter()
        infer.query([free[1]], evidence=rows[0])
        res[label + "_cold_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps({"config": f"{name} single-query latency through the drop-in API", "ve_query_us": res["ve"],
                      "bp_query_us": res["bp"], "ve_new_signature_ms": res["ve_cold_ms"], "bp_new_signature_ms": res["bp_cold_ms"]}), flush=True)


def mixed_ve(name, batch=262144, n_sig=16, specialize=False, n_streams=1):
    """configs[2]: hepar2 / win95pts VE (min-fill order), mixed evidence: 16 observed sets per batch, equal shares."""
    m = px.get_example_model(name)
    ve = VariableElimination(m)
    nodes = sorted(m.nodes())
    per = batch // n_sig
    rng = np.random.default_rng(7)
    q = nodes[len(nodes) // 2]
    jobs = []
    t0 = time.perf_counter()
    for sgn in range(n_sig):
        cand = [v for v in nodes if v != q]
        ev_vars = [cand[i] for i in rng.choice(len(cand), 8, replace=False)]
        _, states = sample_evidence(m, per, 8, seed=100 + sgn, evidence_vars=ev_vars)
        cp = ve._plan([q], ev_vars, True, None)
        if specialize:
            cp.specialize()
        jobs.append((cp, torch.from_numpy(states).cuda(), torch.empty((per, cp.out_elems), dtype=torch.float64, device="cuda")))
    compile_s = time.perf_counter() - t0

    streams = [torch.cuda.Stream() for _ in range(n_streams)] if n_streams > 1 else []
    from pgmpy_b200.engine import MultiRun

    multi = MultiRun(jobs) if n_streams == 0 else None  # n_streams = 0: one call across the C-ABI for all buckets

    def run():
        if multi is not None:
            multi.run()
            return
        # independent launches of 16 384 evidence sets each: round robin onto a few streams (query_batch_mixed does the same)
        if not streams:
            for cp, ev, o in jobs:
                cp.run(ev, out=o)
            return
        cur = torch.cuda.current_stream()
        for st in streams:
            st.wait_stream(cur)
        for j, (cp, ev, o) in enumerate(jobs):
            with torch.cuda.stream(streams[j % len(streams)]):
                cp.run(ev, out=o)
        for st in streams:
            cur.wait_stream(st)

    ms = timed(run)
    alg = sum(cp.plan.algorithmic_bytes(per) for cp, _, _ in jobs)
    print(json.dumps({"config": f"{name} VE single-variable posterior, mixed evidence ({n_sig} signatures x {per} sets)", "batch": per * n_sig,
                      "variant": jobs[0][0].last_variant, "streams": max(1, n_streams), "one_call_for_all_buckets": n_streams == 0,
                      "plan_compile_s": round(compile_s, 2), "ms_per_batch": ms, "evidence_queries_per_sec": per * n_sig / ms * 1e3,
                      "alg_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / PEAK}), flush=True)


def bp_all_marginals(name, batch, k=8, reps=3):
    """configs[3]/[4]: large-table junction-tree all-marginals (pathfinder, diabetes, munin)."""
    m = px.get_example_model(name)
    bp = BeliefPropagation(m)
    ev_vars, states = sample_evidence(m, batch, k, seed=1)
    t0 = time.perf_counter()
    cp = bp.marginals_plan(ev_vars)
    compile_s = time.perf_counter() - t0
    ev = torch.from_numpy(states).cuda()
    out = torch.empty((batch, cp.out_elems), dtype=torch.float64, device="cuda")
    ms = timed(lambda: cp.run(ev, out=out), warm=2, reps=reps)
    alg = cp.plan.algorithmic_bytes(batch)
    print(json.dumps({"config": f"{name} junction-tree all-variable marginals", "batch": batch, "mode": cp.last_mode, "variant": cp.last_variant,
                      "launches": cp.last_launches, "steps": cp.plan.n_steps, "plan_compile_s": round(compile_s, 2), "ms_per_batch": ms,
                      "evidence_queries_per_sec": batch / ms * 1e3, "alg_bytes_per_evidence_set": alg / batch,
                      "alg_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / PEAK,
                      "workspace_GB": cp.workspace_bytes(batch) / 1e9}), flush=True)


def spec_compare(name, batch=131072, k=5):
    """Junction-tree all-marginals of a small model: the default kernel choice against the plan-specialised kernel."""
    m = px.get_example_model(name)
    bp = BeliefPropagation(m)
    ev_vars, states = sample_evidence(m, batch, k, seed=1)
    cp = bp.marginals_plan(ev_vars)
    ev = torch.from_numpy(states).cuda()
    out = torch.empty((batch, cp.out_elems), dtype=torch.float64, device="cuda")
    ms0 = timed(lambda: cp.run(ev, out=out), warm=2, reps=5)
    v0 = cp.last_variant
    ref = out.clone()
    t0 = time.perf_counter()
    info = cp.specialize()
    spec_s = time.perf_counter() - t0
    ms1 = timed(lambda: cp.run(ev, out=out), warm=2, reps=10)
    err = float(((out - ref).abs() / ref.abs().clamp_min(1e-300)).max())
    # the plan variant with the fewest multiply-adds (what marginals_plan(specialize=True) picks)
    t0 = time.perf_counter()
    cp2 = bp.marginals_plan(ev_vars, specialize=True)
    spec2_s = time.perf_counter() - t0
    out2 = torch.empty((batch, cp2.out_elems), dtype=torch.float64, device="cuda")
    ms2 = timed(lambda: cp2.run(ev, out=out2), warm=2, reps=10)
    col = {seg.vars: (seg.out_offset, seg.table.size) for seg in cp.plan.segments}
    err2 = 0.0
    for seg in cp2.plan.segments:
        o, n = col[seg.vars]
        a, b = out2[:, seg.out_offset:seg.out_offset + n], ref[:, o:o + n]
        err2 = max(err2, float(((a - b).abs() / b.abs().clamp_min(1e-300)).max()))
    print(json.dumps({"config": f"{name} junction-tree all-variable marginals, k = {k}", "batch": batch, "steps": cp.plan.n_steps,
                      "distribute": cp.plan.meta.get("distribute"), "default_variant": v0, "ms_default": ms0, "ms_specialized": ms1,
                      "variant": cp.last_variant, "speedup": ms0 / ms1, "evidence_queries_per_sec": batch / ms1 * 1e3,
                      "specialize_s": round(spec_s, 2), "max_rel_diff_vs_default": err,
                      "flops_plan": {"distribute": cp2.plan.meta.get("distribute"), "factorized": bool(cp2.plan.meta.get("factorized")),
                                     "ms": ms2, "variant": cp2.last_variant, "plan_and_specialize_s": round(spec2_s, 2),
                                     "max_rel_diff_vs_default": err2, "spec": cp2.spec_info()}, "spec": info}), flush=True)


if __name__ == "__main__":
    what = sys.argv[1:] or ["alarm_ve", "mixed_ve", "large", "munin"]
    if "latency" in what:
        single_query_latency("alarm")
        single_query_latency("hepar2")
    if "alarm_ve" in what:
        alarm_ve()
    if "mixed_ve" in what:
        mixed_ve("hepar2")
        mixed_ve("win95pts")
    if "mixed_ve_spec" in what:
        mixed_ve("hepar2", specialize=True)
        mixed_ve("win95pts", specialize=True)
        mixed_ve("hepar2", specialize=True, n_streams=0)
        mixed_ve("win95pts", specialize=True, n_streams=0)
        mixed_ve("hepar2", specialize=False, n_streams=0)
        mixed_ve("win95pts", specialize=False, n_streams=0)
    if "mixed_ve_streams" in what:  # measured slower than one stream (host launch bound): kept for the record
        mixed_ve("hepar2", specialize=True, n_streams=4)
        mixed_ve("win95pts", specialize=True, n_streams=4)
    if "spec" in what:
        for nm in ("child", "alarm", "win95pts", "hepar2"):
            spec_compare(nm)
    if "large" in what:
        bp_all_marginals("pathfinder", 16384)
        bp_all_marginals("diabetes", 2048)
    if what and what[0] == "one":
        bp_all_marginals(what[1], int(what[2]), reps=int(what[3]) if len(what) > 3 else 3)
        sys.exit(0)
    if "munin" in what:
        for b in (64, 256, 1024):
            bp_all_marginals("munin", b)
