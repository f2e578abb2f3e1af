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
    ap.add_argument("--fused-kernel", default="auto", choices=["auto", "generic", "tables-smem", "tables-global"])
    ap.add_argument("--step-kernel", default="auto", choices=["auto", "generic", "tile64"])
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunks", type=int, default=0, help="0 = auto (B/8192, at most 16)")
    return ap.parse_args()


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
    """One evidence set: per-variable BeliefPropagation.query exactly as a pgmpy user loops today
    (pgmpy re-initialises the junction tree after every query, so each query re-calibrates).
    A 4th payload item True = the generous variant: calibrate once per evidence set."""
    model_name, ev_vars, row = payload[:3]
    once = len(payload) > 3 and payload[3]
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
        if bp is None or not once:
            bp = O.BP(jt.cliques, jt.edges, [O.Factor(c, p) for c, p in zip(jt.cliques, jt.potentials)])
            bp.calibrate()
        bp.query([v], ev_idx)
        n += 1
    return n


def cpu_baseline(args, ev_vars, states, cores, budget_s):
    """Times the oracle port on as many evidence sets as fit the budget; returns evidence-queries/s."""
    t0 = time.perf_counter()
    done = 0
    marginals = 0
    if cores <= 1:
        i = 0
        while True:
            marginals += _cpu_worker((args.model, ev_vars, states[i % len(states)]))
            done += 1
            i += 1
            if time.perf_counter() - t0 >= budget_s or done >= len(states):
                break
        elapsed = time.perf_counter() - t0
    else:
        import multiprocessing as mp

        ctx = mp.get_context("spawn")
        with ctx.Pool(cores) as pool:
            # warm the per-process caches, then time rounds of `cores` evidence sets
            pool.map(_cpu_worker, [(args.model, ev_vars, states[0])] * cores)
            t0 = time.perf_counter()
            i = 0
            while True:
                rows = [states[(i + j) % len(states)] for j in range(cores)]
                res = pool.map(_cpu_worker, [(args.model, ev_vars, r) for r in rows])
                marginals += sum(res)
                done += cores
                i += cores
                if time.perf_counter() - t0 >= budget_s:
                    break
            elapsed = time.perf_counter() - t0
    return {
        "value": done / elapsed,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"{done} evidence sets ({marginals} single-variable BeliefPropagation.query calls, junction tree "
        f"re-calibrated per query as the reference does) in {elapsed:.1f}s; numpy fp64 oracle/pgm_oracle.py",
        "marginals_per_sec": marginals / elapsed,
    }


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; pgmpy itself is Python and cannot
    travel to the GPU box) on all host cores, same metric/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pgmpy_b200 as px
    from pgmpy_b200.evidence import sample_evidence

    cfg, k = workload_config(args, args.gpus)
    model = px.get_example_model(args.model)
    ev_vars, states = sample_evidence(model, 4096, k, seed=1)
    cores = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 120.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_baseline(args, ev_vars, states, cores, 0.5)
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vals.append(cpu_baseline(args, ev_vars, states, cores, per_step))
    wall = time.perf_counter() - t0
    value = float(np.mean([v["value"] for v in vals]))
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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": vals[-1]["sample"] + f"; each step a {per_step:.1f}s sample"},
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
    cp = bp.marginals_plan(ev_vars)
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
        e2e = {
            "value": (B * n_gpus * args.steps) / float(te.item()),
            "unit": UNIT,
            "h2d_bytes_per_step": int(B * cp.n_ev * 4),
            "d2h_bytes_per_step": int(B * cp.out_elems * itemsize),
            "rank0_cpu_affinity": numa_cpus,
            "how": "CompiledPlan.run_pinned: pinned host evidence -> H2D -> plan -> D2H pinned posteriors every step, batch cut into "
            "chunks over a 3-stream ring so copies overlap kernels; wall clock incl. final sync, max over ranks",
        }

    # ---- final posterior gather over NVLink (the only collective; not on the inference path) ---
    gather = None
    if world > 1:
        full = torch.empty((world * B, cp.out_elems), dtype=cp.torch_dtype, device=dev)
        dist.all_gather_into_tensor(full, out_dev)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(full, out_dev)
        g1.record()
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather = {"op": "nccl all_gather of [B_shard, out_elems] posteriors", "ms": float(tg.item()),
                  "bytes_per_rank": int(B * cp.out_elems * itemsize)}

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            with open(peaks_path) as f:
                peak = float(json.load(f)["hbm_gbs"])
            peak_src = "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        alg_bytes = cp.plan.algorithmic_bytes(B, itemsize)
        per_launch_ms = max_ms / args.steps
        achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
        roofline = {
            "bound": "hbm",
            "kernel": {"generic": "k_plan_fused", "tables-smem": "k_plan_fused2<smem>", "tables-global": "k_plan_fused2<global>"}.get(
                variant, "k_contract_step (sum over the step sequence)"),
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": None,
            "algorithmic_bytes_per_launch": int(alg_bytes),
            "note": "algorithmic bytes = SURVEY 8(d) step formula over the emitted plan (every operand/result counted "
            "once per step regardless of fusion); the fused kernel keeps messages in L1/L2, so frac may exceed 1. "
            "peak: " + peak_src,
        }
        prof = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                with open(prof) as f:
                    roofline["traffic"] = json.load(f).get(f"{args.model}:{mode}:{B}")
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
            "config": dict(cfg, exec_mode=mode, kernel_variant=variant, distribute=cp.plan.meta.get("distribute"), l2=("work tables live in shared memory; " if variant == "tables-smem" else "workspace %.0f MB streamed per step; " % (cp.workspace_bytes(B) / 1e6))
                           + "posteriors written per step %.0f MB; two evidence batches alternate, no L2 flush needed for a kernel whose HBM traffic is write-only output" % (B * cp.out_elems * itemsize / 1e6)),
            "marginals_per_sec": value * len(cp.plan.segments),
            "roofline": roofline,
            "e2e": e2e,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "wall_ms_per_step": 1e3 * t_wall / args.steps,
        }
        if gather:
            line["posterior_gather"] = gather
        if not args.no_cpu_baseline and n_gpus == 1:
            _, cpu_states = sample_evidence(model, 64, k, seed=1, evidence_vars=ev_vars)
            line["cpu_baseline"] = cpu_baseline(args, ev_vars, cpu_states, 1, args.cpu_seconds)
            # for transparency: the same loop if the junction tree were calibrated only once per evidence set
            t0 = time.perf_counter()
            n_once = 0
            while time.perf_counter() - t0 < max(2.0, args.cpu_seconds / 5):
                _cpu_worker((args.model, ev_vars, cpu_states[n_once % len(cpu_states)], True))
                n_once += 1
            line["cpu_baseline"]["value_if_calibrated_once_per_evidence_set"] = n_once / (time.perf_counter() - t0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


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
