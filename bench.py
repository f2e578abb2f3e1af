#!/usr/bin/env python
"""bench.py — evidence-query posteriors/sec for batched exact inference on B200.

Workload (BASELINE.json configs[1]): alarm (37 nodes), BeliefPropagation-mode junction-tree inference,
posterior marginals of ALL unobserved variables for each evidence set; 1,048,576 forward-sampled evidence
sets sharded over 8 B200s = 131,072 per GPU (weak scaling: per-GPU batch fixed), fp64, k = 5 observed
variables. A "step" = one pass of the whole plan over one per-GPU batch.

    python bench.py --gpus N --steps K --warmup W           # N>1: launched by torchrun, one rank per GPU
    python bench.py --impl reference ...                    # the reference's CPU algorithm (oracle port)

One JSON line on rank 0. See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

METRIC = "evidence_queries_per_sec"
UNIT = "evidence-queries/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="alarm")
    ap.add_argument("--batch-per-gpu", type=int, default=131072)
    ap.add_argument("--n-evidence", type=int, default=None)
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32"])
    ap.add_argument("--mode", default="auto", choices=["auto", "fused", "stepwise"])
    ap.add_argument("--fused-warps", type=int, default=0)
    ap.add_argument("--fused-kernel", default="auto", choices=["auto", "generic", "tables-smem", "tables-global", "specialized"])
    ap.add_argument("--no-specialize", action="store_true",
                    help="do not build the plan-specialised kernel (pgx_plan_specialize): time the table-driven kernel")
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "generic", "tile64"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="0 = auto (B/8192, at most 16)")
    ap.add_argument("--no-configs", action="store_true", help="skip the extra BASELINE configs (munin / diabetes / pathfinder)")
    ap.add_argument("--reference-kind", default="auto", choices=["auto", "reference", "port"],
                    help="--impl reference: real pgmpy from oracle/_ref (or /root/reference) | the numpy port")
    return ap.parse_args()


# BASELINE.json configs[3] / configs[4]: the HBM-resident junction trees, per-GPU batch, k = 8 evidence variables
EXTRA_CONFIGS = (("munin", 256), ("diabetes", 2048), ("pathfinder", 16384))


def workload_config(args, n_gpus):
    k = args.n_evidence if args.n_evidence is not None else (5 if args.model == "alarm" else 8)
    return {
        "workload": f"{args.model} junction-tree BeliefPropagation, all-variable posterior marginals per evidence set",
        "model": args.model,
        "evidence_vars_per_set": k,
        "batch_per_gpu": args.batch_per_gpu,
        "global_batch": args.batch_per_gpu * n_gpus,
        "evidence": "forward-sampled (ancestral) assignments, k revealed variables, seed 1",
        "parallelism": f"evidence batch sharded over {n_gpus} GPU(s), no collective on the inference path",
    }, k


# ---------------------------------------------------------------------------------------------
# clocks sampling (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = (
        "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, first=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[first:]:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(mx)) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference algorithm, timed on a bounded sample
# ---------------------------------------------------------------------------------------------
def _cpu_worker(payload):
    """One evidence set on the numpy port of the reference algorithm (oracle/pgm_oracle.py): calibrate the junction tree
    once, then one BeliefPropagation.query([v], e) per unobserved variable. A 4th payload item True = additionally
    re-calibrate before EVERY query, which is what the reference's public query() does (it re-initialises itself after
    each call, ExactInference.py:1215) — reported as a secondary figure only."""
    model_name, ev_vars, row = payload[:3]
    recalibrate = len(payload) > 3 and payload[3]
    import pgmpy_b200 as px
    from oracle import pgm_oracle as O
    from pgmpy_b200.planner import JTStructure

    cache = _cpu_worker.__dict__.setdefault("cache", {})
    if model_name not in cache:
        m = px.get_example_model(model_name)
        jt = JTStructure.from_model(m)
        cache[model_name] = (m, jt)
    m, jt = cache[model_name]
    ev_idx = {v: int(s) for v, s in zip(ev_vars, row)}
    free = [v for c in jt.cliques for v in c]
    seen = set()
    n = 0
    bp = None
    for v in free:
        if v in seen or v in ev_idx:
            continue
        seen.add(v)
        if bp is None or recalibrate:
            bp = O.BP(jt.cliques, jt.edges, [O.Factor(c, p) for c, p in zip(jt.cliques, jt.potentials)])
            bp.calibrate()
        bp.query([v], ev_idx)
        n += 1
    return n


def _ref_worker(payload):
    """One evidence set on the UNMODIFIED reference (pgmpy 1.0.0 from oracle/_ref or /root/reference), public API:
    BeliefPropagation(<our min-fill junction tree as a pgmpy JunctionTree>), calibrate() once, then
    query([v], evidence) for every unobserved variable. pgmpy's query() re-initialises the object after each call
    (ExactInference.py:1215); the calibrated beliefs (public attributes) are put back so that an evidence set costs ONE
    calibration, like the GPU arm. (pgmpy's own junction-tree builder needs 25-83 s per calibration on alarm,
    SURVEY.md fact 5, so the tree is ours.)"""
    model_name, ev_vars, row = payload[:3]
    cache = _ref_worker.__dict__.setdefault("cache", {})
    if model_name not in cache:
        from oracle.ref_loader import load_reference

        load_reference()
        from pgmpy.factors.discrete import DiscreteFactor as RefDF
        from pgmpy.inference import BeliefPropagation as RefBP
        from pgmpy.models import JunctionTree as RefJT

        import pgmpy_b200 as px
        from pgmpy_b200.planner import JTStructure

        m = px.get_example_model(model_name)
        jt = JTStructure.from_model(m)
        rjt = RefJT()
        for c in jt.cliques:
            rjt.add_node(c)
        for a, b in jt.edges:
            rjt.add_edge(jt.cliques[a], jt.cliques[b])
        for c, p in zip(jt.cliques, jt.potentials):
            rjt.add_factors(RefDF(list(c), [jt.card[v] for v in c], p, state_names={v: m.states[v] for v in c}))
        cache[model_name] = (m, jt, RefBP(rjt))
    m, jt, bp = cache[model_name]
    ev = {v: m.states[v][int(s)] for v, s in zip(ev_vars, row)}
    seen = set()
    n = 0
    with np.errstate(all="ignore"):
        bp.calibrate()
        cb, sb = bp.clique_beliefs, bp.sepset_beliefs
        for c in jt.cliques:
            for v in c:
                if v in seen or v in ev:
                    continue
                seen.add(v)
                bp.clique_beliefs, bp.sepset_beliefs = cb, sb
                bp.query([v], evidence=ev, show_progress=False)
                n += 1
    return n


def reference_kind(args):
    if getattr(args, "reference_kind", "auto") == "port":
        return "port"
    try:
        from oracle.ref_loader import reference_available

        if reference_available():
            return "reference"
    except Exception:
        pass
    if getattr(args, "reference_kind", "auto") == "reference":
        raise RuntimeError("the reference is not installed (oracle/_ref missing: run __graft_entry__.build() in the build container)")
    return "port"


def cpu_baseline(args, ev_vars, states, cores, budget_s, kind="port", recalibrate=False):
    """Times the CPU arm on as many evidence sets as fit the budget; returns evidence-queries/s."""
    worker = _ref_worker if kind == "reference" else _cpu_worker
    extra = (True,) if (recalibrate and kind == "port") else ()
    t0 = time.perf_counter()
    done = 0
    marginals = 0
    if cores <= 1:
        worker((args.model, ev_vars, states[0]) + extra)  # warm the caches (model load, junction tree), untimed
        t0 = time.perf_counter()
        i = 0
        while True:
            marginals += worker((args.model, ev_vars, states[i % len(states)]) + extra)
            done += 1
            i += 1
            if time.perf_counter() - t0 >= budget_s or done >= len(states):
                break
        elapsed = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        pool = cpu_baseline.__dict__.get("pool")
        if pool is None or cpu_baseline.__dict__.get("pool_key") != (cores, kind):
            if pool is not None:
                pool.terminate()
            pool = mp.get_context("spawn").Pool(cores)
            cpu_baseline.pool, cpu_baseline.pool_key = pool, (cores, kind)
            # warm the per-process caches (import, model load, junction tree), untimed
            pool.map(worker, [(args.model, ev_vars, states[0]) + extra] * cores)
        t0 = time.perf_counter()
        i = 0
        while True:
            rows = [states[(i + j) % len(states)] for j in range(cores)]
            res = pool.map(worker, [(args.model, ev_vars, r) + extra for r in rows])
            marginals += sum(res)
            done += cores
            i += cores
            if time.perf_counter() - t0 >= budget_s:
                break
        elapsed = time.perf_counter() - t0
    what = ("pgmpy 1.0.0 BeliefPropagation on our min-fill junction tree: calibrate() once per evidence set, then query([v], e) per "
            "unobserved variable (public API, numpy backend, fp64)" if kind == "reference" else
            "numpy port of the reference algorithm (oracle/pgm_oracle.py): junction tree calibrated "
            + ("before every query, as the reference's query() re-initialises itself" if recalibrate else "once per evidence set")
            + ", then one query per unobserved variable")
    return {
        "value": done / elapsed,
        "unit": UNIT,
        "cores": cores,
        "kind": kind,
        "sample": f"{done} evidence sets ({marginals} single-variable queries) in {elapsed:.1f}s; {what}",
        "marginals_per_sec": marginals / elapsed,
    }


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on all host cores, same metric/config.
    Real pgmpy (kind "reference") when oracle/_ref travelled with the snapshot, else the numpy port (kind "port")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence

    cfg, k = workload_config(args, args.gpus)
    kind = reference_kind(args)
    model = px.get_example_model(args.model)
    ev_vars, states = sample_evidence(model, 4096, k, seed=1)
    cores = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_baseline(args, ev_vars, states, cores, 0.5, kind)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_baseline(args, ev_vars, states, cores, per_step, kind))
    wall = time.perf_counter() - t0
    value = float(np.mean([v["value"] for v in vals]))
    base = {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": vals[-1]["sample"] + f"; each step a >= {per_step:.1f}s sample"}
    if kind == "reference":
        # the numpy port on the same loop, for continuity with round 1 (which had only the port)
        base["port_value"] = cpu_baseline(args, ev_vars, states, cores, min(per_step, 5.0), "port")["value"]
    pool = cpu_baseline.__dict__.get("pool")
    if pool is not None:
        pool.terminate()
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * wall / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": cfg,
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence
    from pgmpy_b200.inference import BeliefPropagation

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a GPU: pgmpy_b200 has no CPU execution path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from pgmpy_b200.distributed import bind_process_to_gpu_numa

    numa_cpus = bind_process_to_gpu_numa(local_rank) if world > 1 else None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # NCCL's banner/log lines go to stderr: stdout = one JSON line
        dist.init_process_group("nccl", device_id=dev)
    n_gpus = world
    cfg, k = workload_config(args, n_gpus)
    B = args.batch_per_gpu

    model = px.get_example_model(args.model)
    bp = BeliefPropagation(model, dtype=args.dtype)
    # same evidence-variable set on every rank; each rank draws its own shard of evidence sets
    ev_vars, _ = sample_evidence(model, 1, k, seed=1)
    spec = None
    want_spec = not args.no_specialize and args.mode != "stepwise" and args.fused_kernel in ("auto", "specialized")
    # one-off per evidence signature (NVRTC, about a second): the plan variant with the fewest multiply-adds, as
    # straight-line sm_100a code. Plans the generator refuses (max-product, soft evidence, too large) or a box without
    # libnvrtc keep the table-driven kernel, and the line says so.
    t_spec = time.perf_counter()
    cp = bp.marginals_plan(ev_vars, specialize=want_spec)
    if want_spec:
        spec = cp.spec_info()
        spec["plan_and_specialize_s"] = round(time.perf_counter() - t_spec, 2)
        if not spec["specialized"]:
            try:
                cp.specialize()
            except Exception as exc:
                if args.fused_kernel == "specialized":
                    raise
                spec["why"] = str(exc)[:200]
    cp.set_mode(args.mode, args.fused_warps, args.fused_kernel, args.step_kernel)
    n_batches = 2
    shards = []
    for j in range(n_batches):
        _, st = sample_evidence(model, B, k, seed=1000 * (j + 1) + rank, evidence_vars=ev_vars)
        shards.append(st)
    ev_dev = [torch.from_numpy(s).to(dev) for s in shards]
    out_dev = torch.empty((B, cp.out_elems), dtype=cp.torch_dtype, device=dev)
    itemsize = 8 if args.dtype == "float64" else 4

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") -------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        cp.run(ev_dev[i % n_batches], out=out_dev)
    barrier()
    if rank == 0:
        # keep the GPU busy (untimed) until nvidia-smi has produced its first samples
        t_wait = time.perf_counter()
        while len(sampler.lines) < 2 and time.perf_counter() - t_wait < 3.0:
            cp.run(ev_dev[0], out=out_dev)
            torch.cuda.synchronize()
    barrier()
    n_before = len(sampler.lines)
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall = time.perf_counter()
    start.record()
    for i in range(args.steps):
        cp.run(ev_dev[i % n_batches], out=out_dev)
    stop.record()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = start.elapsed_time(stop)
    launches = cp.last_launches * args.steps
    mode = cp.last_mode
    variant = cp.last_variant
    if rank == 0 and len(sampler.lines) - n_before < 2:
        # very short timed region: take a few more samples under the same load (untimed)
        t_wait = time.perf_counter()
        while len(sampler.lines) - n_before < 3 and time.perf_counter() - t_wait < 2.0:
            cp.run(ev_dev[0], out=out_dev)
            torch.cuda.synchronize()
    clocks = sampler.stop(n_before) if rank == 0 else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = (B * n_gpus * args.steps) / (max_ms * 1e-3)

    # ---- end to end through the public API with HOST buffers ("e2e") ---------------------------
    e2e = None
    if not args.no_e2e:
        ev_pin = [torch.from_numpy(s).pin_memory() for s in shards]
        out_pin = torch.empty((B, cp.out_elems), dtype=cp.torch_dtype).pin_memory()
        for i in range(max(2, min(args.warmup, 3))):
            cp.run_pinned(ev_pin[i % n_batches], out_pin, args.e2e_chunks)
        barrier()
        e0 = time.perf_counter()
        for i in range(args.steps):
            cp.run_pinned(ev_pin[i % n_batches], out_pin, args.e2e_chunks)
        barrier()
        e_s = time.perf_counter() - e0
        te = torch.tensor([e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # the ceiling of that path: a bare pinned device->host copy of the same posterior buffer, every rank at once
        for _ in range(2):
            out_pin.copy_(out_dev, non_blocking=True)
        barrier()
        c0 = time.perf_counter()
        n_copy = 10
        for _ in range(n_copy):
            out_pin.copy_(out_dev, non_blocking=True)
        barrier()
        tc = torch.tensor([time.perf_counter() - c0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        d2h_bytes = int(B * cp.out_elems * itemsize)
        d2h_gbs = d2h_bytes * n_copy / float(tc.item()) / 1e9
        e2e = {
            "value": (B * n_gpus * args.steps) / float(te.item()),
            "unit": UNIT,
            "h2d_bytes_per_step": int(B * cp.n_ev * 4),
            "d2h_bytes_per_step": d2h_bytes,
            "d2h_ceiling_gbs_per_gpu": d2h_gbs,
            "d2h_ceiling_value": B * n_gpus / (d2h_bytes / (d2h_gbs * 1e9)),
            "rank0_cpu_affinity": numa_cpus,
            "how": "CompiledPlan.run_pinned: pinned host evidence -> H2D -> plan -> D2H pinned posteriors every step, batch cut into "
            "chunks over a 3-stream ring so copies overlap kernels; wall clock incl. final sync, max over ranks. "
            "d2h_ceiling_*: a bare pinned D2H copy loop of the same posterior buffer on every rank at once (slowest rank) and the "
            "evidence-queries/s it would allow: what the host side of this box can take, whatever the kernels do",
        }

    # ---- final posterior gather over NVLink (the only collective; not on the inference path) ---
    gather = None
    if world > 1:
        from pgmpy_b200.distributed import gather_posteriors_to_root

        gather_posteriors_to_root(out_dev)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_g = 5
        g0.record()
        for i in range(n_g):
            cp.run(ev_dev[i % n_batches], out=out_dev)
            gather_posteriors_to_root(out_dev)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1) / n_g], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"op": "NCCL gather of the [B_shard, out_elems] posteriors to rank 0 after every step (device resident)",
                  "ms_per_step_with_gather": float(tg.item()), "bytes_per_rank": int(B * cp.out_elems * itemsize),
                  "value_with_gather": B * n_gpus / (float(tg.item()) * 1e-3)}

    # ---- the HBM-resident BASELINE configs (munin / diabetes / pathfinder), same timing rules ---
    extra = {}
    if not args.no_configs and args.model == "alarm":
        for name, bsz in EXTRA_CONFIGS:
            extra[f"{name}_B{bsz}"] = bench_config(name, bsz, args, world, rank, dev, barrier)

    if rank == 0:
        peak, peak_src = measured_peak()
        alg_bytes = cp.plan.algorithmic_bytes(B, itemsize)
        per_launch_ms = max_ms / args.steps
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        fused = variant in ("generic", "tables-smem", "tables-global", "specialized")
        roofline = {
            # the whole-plan kernel keeps its work tables in shared memory: HBM sees evidence in and posteriors out only
            # (`traffic`), so the binding resource is the SM's own load/store path and the per-level barriers; the
            # GB/s figure below is the notional one of SURVEY 8(d) (every operand counted as if it moved)
            "bound": "l1-lsu+barrier" if fused else "hbm",
            "kernel": {"generic": "k_plan_fused", "tables-smem": "k_plan_fused2<smem>", "tables-global": "k_plan_fused2<global>",
                       "specialized": "k_plan_spec (plan-specialised, NVRTC sm_100a)"}.get(
                variant, "k_contract_tile32 / k_contract_mm (sum over the launch sequence)"),
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": None,
            "algorithmic_bytes_per_launch": int(alg_bytes),
            "note": "algorithmic bytes = SURVEY 8(d) step formula over the emitted plan (every operand/result counted "
            "once per step regardless of fusion). peak: " + peak_src,
        }
        if variant == "specialized":
            # the specialised kernel keeps every message in registers: the only bytes that exist outside the SM are the
            # evidence rows in and the posterior rows out, and that write stream is the roofline that bounds it (a perfect
            # kernel would take io_bytes / peak). `achieved`/`frac` are therefore quoted on those I/O bytes; the SURVEY 8(d)
            # step-formula figure (every operand counted as if it moved, > peak by construction here) stays beside it
            io_bytes = B * cp.n_ev * 4 + B * cp.out_elems * itemsize
            roofline.update({
                "bound": "hbm",
                "achieved": io_bytes / (per_launch_ms * 1e-3) / 1e9,
                "frac": io_bytes / (per_launch_ms * 1e-3) / 1e9 / peak,
                "io_bytes_per_launch": int(io_bytes),
                "survey_8d_step_formula": {"bytes_per_launch": int(alg_bytes), "GBps": achieved, "frac": achieved / peak},
                "limiter": "half way to HBM: the compute part (0.028 of 0.030 ms; persistent CTAs, the TMA bulk store of a row drains "
                "under the next row's compute) keeps the fp64 pipe 56 % busy with 8 warps per SM at 255 registers; ncu: issue-active "
                "45 %, stalls wait / selected / long_scoreboard",
                "note": "achieved = (evidence in + posteriors out) / kernel time; traffic = ncu DRAM bytes of one launch (the rest of "
                "the 95 MB of posteriors is still in the 126 MB L2 when the kernel ends). peak: " + peak_src,
            })
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    tr = json.load(f)
                key = f"{args.model}:{'specialized' if variant == 'specialized' else mode}:{B}"
                roofline["traffic"] = tr.get(key)
                roofline["ncu"] = tr.get(key + ":ncu")
            except Exception:
                pass
        line = {
            "metric": METRIC,
            "value": value,
            "unit": UNIT,
            "n_gpus": n_gpus,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": per_launch_ms,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f64" if args.dtype == "float64" else "f32",
            "data": "synthetic",
            "config": cfg,
            "engine": {"exec_mode": mode, "kernel_variant": variant, "distribute": cp.plan.meta.get("distribute"),
                       "factorized_potentials": bool(cp.plan.meta.get("factorized")),
                       "specialized_kernel": spec,
                       "l2": ("work tables live in shared memory; " if variant in ("tables-smem", "specialized") else "workspace %.0f MB streamed per step; " % (cp.workspace_bytes(B) / 1e6))
                       + "posteriors written per step %.0f MB; two evidence batches alternate, no L2 flush needed for a kernel whose HBM traffic is write-only output" % (B * cp.out_elems * itemsize / 1e6)},
            "marginals_per_sec": value * len(cp.plan.segments),
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
        }
        if extra:
            line["configs"] = extra
        if gather:
            line["posterior_gather"] = gather
        if not args.no_cpu_baseline and n_gpus == 1:
            kind = reference_kind(args)
            _, cpu_states = sample_evidence(model, 64, k, seed=1, evidence_vars=ev_vars)
            line["cpu_baseline"] = cpu_baseline(args, ev_vars, cpu_states, 1, args.cpu_seconds, kind)
            if kind == "reference":
                line["cpu_baseline"]["port_value"] = cpu_baseline(args, ev_vars, cpu_states, 1, max(2.0, args.cpu_seconds / 5), "port")["value"]
            # secondary: what a user of the reference's public query() pays when nothing is put back between queries
            line["cpu_baseline"]["port_value_recalibrating_before_every_query"] = cpu_baseline(
                args, ev_vars, cpu_states, 1, max(2.0, args.cpu_seconds / 5), "port", recalibrate=True)["value"]
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measured_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


FP64_PEAK_TFLOPS = 33.8  # measured DFMA rate of this pool's B200 (tools/microbench/dmma_vs_dfma.cu, profiles/r02_dmma_vs_dfma.json)


def bench_config(name, bsz, args, world, rank, dev, barrier):
    """One of the HBM-resident BASELINE configs: junction-tree all-variable marginals, `bsz` evidence sets per GPU
    (k = 8 forward-sampled evidence variables), device resident, CUDA events, max over ranks. Two evidence batches
    alternate and the work tables of one pass (GBs) exceed L2 many times over, so no L2 flush is needed."""
    import torch
    import torch.distributed as dist

    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence
    from pgmpy_b200.inference import BeliefPropagation

    t0 = time.perf_counter()
    model = px.get_example_model(name)
    bp = BeliefPropagation(model, dtype=args.dtype)
    ev_vars, _ = sample_evidence(model, 1, 8, seed=1)
    cp = bp.marginals_plan(ev_vars)
    compile_s = time.perf_counter() - t0
    evs = []
    for j in range(2):
        _, st = sample_evidence(model, bsz, 8, seed=1000 * (j + 1) + rank, evidence_vars=ev_vars)
        evs.append(torch.from_numpy(st).to(dev))
    out = torch.empty((bsz, cp.out_elems), dtype=cp.torch_dtype, device=dev)
    reps = 4
    for i in range(4):  # (both evidence buffers twice: the launch sequence is captured as a CUDA graph on the second use)
        cp.run(evs[i % 2], out=out)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        cp.run(evs[i % 2], out=out)
    b.record()
    barrier()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    itemsize = 8 if args.dtype == "float64" else 4
    alg = cp.plan.algorithmic_bytes(bsz, itemsize)
    flops = 2 * cp.plan.flops(bsz)  # one multiply + one add per operand load, an upper bound
    peak, _ = measured_peak()
    res = {
        "workload": f"{name} junction-tree BeliefPropagation, all-variable posterior marginals, {bsz} evidence sets per GPU, k = 8",
        "ms_per_batch": ms,
        "evidence_queries_per_sec": bsz * world / (ms * 1e-3),
        "algorithmic_bytes_per_evidence_set": alg / bsz,
        "algorithmic_GBps": alg / ms / 1e6,
        "frac_of_hbm_peak": alg / ms / 1e6 / peak,
        "operand_loads_per_evidence_set": cp.plan.operand_loads(),
        "fp64_tflops": flops / ms / 1e9,
        "frac_of_fp64_peak": flops / ms / 1e9 / FP64_PEAK_TFLOPS,
        "launches": int(cp.last_launches),
        "steps": cp.plan.n_steps,
        "staged_gemm_steps": int(cp.last_staged_steps),
        "plan": {"distribute": cp.plan.meta.get("distribute"), "factorized_potentials": bool(cp.plan.meta.get("factorized"))},
        "plan_compile_s": round(compile_s, 2),
        "workspace_GB": cp.workspace_bytes(bsz) / 1e9,
    }
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            res["dram_traffic_bytes_per_batch_ncu"] = json.load(f).get(f"{name}:stepwise:{bsz}")  # real HBM bytes, one ncu pass
    except Exception:
        res["dram_traffic_bytes_per_batch_ncu"] = None
    r1 = R01_PLAN_BYTES.get(name)
    if r1:
        # continuity with round 1, whose plans moved more bytes for the same posteriors: the same time against THAT count
        res["round1_plan_bytes_per_evidence_set"] = r1
        res["frac_of_hbm_peak_on_round1_bytes"] = r1 * bsz / ms / 1e6 / peak
    del cp, bp, out, evs
    torch.cuda.empty_cache()
    return res


# algorithmic bytes per evidence set of the round-1 plans of the same queries (profiles/r01_baseline_configs.jsonl)
R01_PLAN_BYTES = {"munin": 845359612, "diabetes": 83620128, "pathfinder": 4218862}


def main():
    args = parse_args()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # launched as plain `python bench.py --gpus N`: re-exec under torchrun, one rank per GPU
        import socket

        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
                                   "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:])
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
