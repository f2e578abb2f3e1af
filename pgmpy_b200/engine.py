"""Device-side plan objects: upload a compiled Plan once, run it over batches of evidence sets.

PyTorch is only the buffer carrier here (device memory, streams); all arithmetic happens in the
hand-written kernels of libpgx.so. Without CUDA every entry point raises — no CPU fallback.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional

import numpy as np

from . import _native as N
from .plan import Plan


def _torch():
    import torch

    return torch


def require_cuda():
    torch = _torch()
    if not torch.cuda.is_available():
        raise RuntimeError("pgmpy_b200 needs a CUDA device (B200); there is no CPU execution path")
    return torch


class MultiRun:
    """A fixed list of (plan, evidence tensor, output tensor) jobs enqueued by ONE call across the C-ABI
    (pgx_run_batch_multi): the buckets of a mixed-evidence batch. Building the argument arrays once and replaying them
    takes the per-bucket host path (~15 us of Python + ctypes each) out of the loop."""

    def __init__(self, jobs):
        torch = require_cuda()
        self.jobs = list(jobs)
        n = len(self.jobs)
        if n == 0:
            raise ValueError("no jobs")
        self.lib = self.jobs[0][0].lib
        self.device = self.jobs[0][0].device
        self._keep = []
        plans, evs, outs, wss, wsb, bs = [], [], [], [], [], []
        for cp, ev, out in self.jobs:
            if cp.device != self.device:
                raise ValueError("all plans of a MultiRun must live on one device")
            if cp.plan.in_elems:
                raise ValueError("plans with soft-evidence inputs cannot be part of a MultiRun")
            B = int(out.shape[0])
            if cp.n_ev and (ev is None or ev.dtype != torch.int32 or not ev.is_cuda or not ev.is_contiguous() or tuple(ev.shape) != (B, cp.n_ev)):
                raise ValueError(f"evidence must be a contiguous int32 CUDA tensor [{B}, {cp.n_ev}]")
            if out.dtype != cp.torch_dtype or not out.is_contiguous() or tuple(out.shape) != (B, cp.out_elems) or not out.is_cuda:
                raise ValueError("out has the wrong dtype/shape")
            need = cp.workspace_bytes(B)
            ws = torch.empty(max(need, 1), dtype=torch.uint8, device=self.device)
            self._keep.append(ws)
            plans.append(cp.handle.value)
            evs.append(ev.data_ptr() if cp.n_ev else 0)
            outs.append(out.data_ptr())
            wss.append(ws.data_ptr())
            wsb.append(ws.numel())
            bs.append(B)
        self.n = n
        self._plans = (C.c_void_p * n)(*plans)
        self._evs = (C.c_void_p * n)(*evs)
        self._outs = (C.c_void_p * n)(*outs)
        self._wss = (C.c_void_p * n)(*wss)
        self._wsb = (C.c_size_t * n)(*wsb)
        self._bs = (C.c_int64 * n)(*bs)

    def run(self):
        """Enqueue every job on the current stream (asynchronous)."""
        torch = _torch()
        guard = contextlib.nullcontext() if torch.cuda.current_device() == self.device.index else torch.cuda.device(self.device)
        with guard:
            stream = torch.cuda.current_stream(self.device).cuda_stream
            N.check(self.lib.pgx_run_batch_multi(
                self.n, self._plans, self._evs, self._outs, self._wss, self._wsb, self._bs, C.c_void_p(stream)))


class CompiledPlan:
    """A Plan resident on one GPU. `dtype`: "float64" (default, 1e-12 parity target) or "float32"."""

    # A plan that has served this many evidence sets is specialised automatically on its next run (None: never).
    # NVRTC + load cost about a second; the specialised kernel saves ~3.4 ns per evidence set on alarm, so the switch
    # pays for itself after a few 10^8 sets. Callers who know their volume call specialize() up front (bench.py does).
    AUTO_SPECIALIZE_SETS = 1 << 27

    def __init__(self, plan: Plan, dtype: str = "float64", device: Optional[int] = None):
        torch = require_cuda()
        self.lib = N.load()
        self.plan = plan
        self.dtype_name = dtype
        self.torch_dtype = {"float64": torch.float64, "float32": torch.float32}[dtype]
        self.pgx_dtype = N.PGX_F64 if dtype == "float64" else N.PGX_F32
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        with torch.cuda.device(self.device):
            self.blob = torch.from_numpy(plan.const_blob).to(self.device, dtype=self.torch_dtype)
            pool = np.ascontiguousarray(plan.pool, dtype=np.int32)
            desc = N.PlanDesc(
                1,
                self.pgx_dtype,
                pool.ctypes.data_as(C.POINTER(C.c_int32)),
                pool.size,
                C.c_void_p(self.blob.data_ptr()),
                self.blob.numel(),
            )
            handle = C.c_void_p()
            N.check(self.lib.pgx_plan_create(C.byref(desc), C.byref(handle)))
        self.handle = handle
        self.n_ev = len(plan.ev_vars)
        self.out_elems = plan.out_elems
        self._ws = None
        self._sets_seen = 0
        self._spec_tried = False

    def __del__(self):
        h = getattr(self, "handle", None)
        if h:
            try:
                self.lib.pgx_plan_destroy(h)
            except Exception:
                pass
            self.handle = None

    # ---- options / info ----------------------------------------------------------------------
    def set_mode(self, mode: str = "auto", fused_warps: int = 0, fused_kernel: str = "auto", step_kernel: str = "auto"):
        """mode: auto | stepwise | fused.  fused_kernel: auto | generic | tables-smem | tables-global | specialized.
        step_kernel: auto (tile-cooperative, 32-bit addressing) | generic | tile64."""
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_STEP_KERNEL, {"auto": 0, "generic": 1, "tile64": 2}[step_kernel]))
        m = {"auto": N.MODE_AUTO, "stepwise": N.MODE_STEPWISE, "fused": N.MODE_FUSED}[mode]
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_MODE, m))
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_FUSED_WARPS, fused_warps))
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_FUSED_KERNEL, N.FUSED_KERNELS[fused_kernel]))

    def specialize(self) -> dict:
        """Build the plan-specialised kernel (pgx_plan_specialize): the plan's fixed multiply-add list as straight-line
        sm_100a code, compiled with NVRTC (seconds). Later run() calls use it unless set_mode asks for another kernel.
        Raises PgxError when the plan cannot be specialised (divide / max steps, soft-evidence inputs, too many product
        terms, no libnvrtc). Returns the kernel's statistics."""
        torch = _torch()
        self._spec_tried = True
        with torch.cuda.device(self.device):
            N.check(self.lib.pgx_plan_specialize(self.handle))
        return self.spec_info()

    def spec_info(self) -> dict:
        return {"specialized": bool(self.info(N.INFO_SPECIALIZED)), "registers": self.info(N.INFO_SPEC_REGS),
                "smem_bytes": self.info(N.INFO_SPEC_SMEM), "compile_ms": self.info(N.INFO_SPEC_COMPILE_MS),
                "loads_per_row": self.info(N.INFO_SPEC_LOADS), "fp_instr_per_row": self.info(N.INFO_SPEC_FLOPS)}

    def info(self, what: int) -> int:
        v = C.c_int64()
        N.check(self.lib.pgx_plan_get_info(self.handle, what, C.byref(v)))
        return v.value

    @property
    def last_launches(self) -> int:
        return self.info(N.INFO_LAST_LAUNCHES)

    @property
    def last_mode(self) -> str:
        return {N.MODE_STEPWISE: "stepwise", N.MODE_FUSED: "fused", 0: "none"}[self.info(N.INFO_LAST_MODE)]

    @property
    def last_variant(self) -> str:
        v = self.info(N.INFO_LAST_VARIANT)
        return {0: "stepwise", 1: "generic", 2: "tables-smem", 3: "tables-global", 4: "specialized"}[v]

    def set_stage(self, which=1):
        """Matrix-product-shaped two-operand steps: 1/True the pipelined shared-memory-staged tile kernel k_contract_mm
        (default), 0/False the streaming kernel only."""
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_STAGE, int(which)))

    def set_mma(self, enabled: bool = True):
        """fp64 tensor cores (DMMA) for k_contract_mm steps with a batch-invariant first operand (default on)."""
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_MMA, 1 if enabled else 0))

    def set_tc32(self, enabled: bool = True):
        """fp32 mode: CPT-times-message steps on the tcgen05 tensor cores (TF32x3, TMEM accumulators; default on)."""
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_TC32, 1 if enabled else 0))

    @property
    def last_tc_steps(self) -> int:
        return self.info(N.INFO_LAST_TC_STEPS)

    def set_graph(self, enabled: bool = True):
        N.check(self.lib.pgx_plan_set_option(self.handle, N.OPT_USE_GRAPH, 1 if enabled else 0))

    @property
    def last_staged_steps(self) -> int:
        return self.info(N.INFO_LAST_STAGED_STEPS)

    @property
    def last_graph(self) -> bool:
        return bool(self.info(N.INFO_LAST_GRAPH))

    def workspace_bytes(self, batch: int) -> int:
        return int(self.lib.pgx_workspace_bytes(self.handle, batch))

    # ---- execution ---------------------------------------------------------------------------
    # Largest workspace a single launch sequence may ask for; bigger batches are processed in row tiles
    # (munin needs 169 MB of work tables per evidence set: 1024 sets would want 173 GB at once).
    MAX_WORKSPACE_BYTES = 31 << 30  # also keeps every element index of the 32-bit-addressed step kernel below 2^32
    MAX_ROWS_PER_PASS = 1 << 20  # the stepwise grid covers at most 65 535 tiles of 32 evidence sets

    def run(self, ev_states, out=None, workspace=None, soft=None):
        """ev_states: int32 CUDA tensor [B, n_ev] (or [B, 0] / None with B given by `out`).
        soft: CUDA tensor [B, in_elems] of the plan dtype when the plan has batch-dependent input tables (soft
        evidence likelihoods; layout in plan.inputs), else None.
        Returns out: [B, out_elems] CUDA tensor of the plan dtype. Asynchronous on the current stream."""
        torch = _torch()
        in_elems = self.plan.in_elems
        if in_elems and soft is None:
            raise ValueError("this plan has soft-evidence input tables: pass `soft` [B, in_elems]")
        if soft is not None:
            if not in_elems:
                raise ValueError("this plan has no soft-evidence input tables")
            if soft.dtype != self.torch_dtype or not soft.is_cuda or not soft.is_contiguous() or soft.dim() != 2 or soft.shape[1] != in_elems:
                raise ValueError(f"soft must be a contiguous {self.dtype_name} CUDA tensor [B, {in_elems}]")
            if ev_states is not None and soft.shape[0] != ev_states.shape[0]:
                raise ValueError("soft and ev_states disagree on the batch size")
        if workspace is None:
            B_all = ev_states.shape[0] if ev_states is not None else (out.shape[0] if out is not None else (soft.shape[0] if soft is not None else 0))
            if B_all > 32 and (self.workspace_bytes(B_all) > self.MAX_WORKSPACE_BYTES or B_all > self.MAX_ROWS_PER_PASS):
                per_set = self.workspace_bytes(32) // 32
                tile = max(32, int(self.MAX_WORKSPACE_BYTES // max(per_set, 1)) // 32 * 32)
                tile = min(tile, self.MAX_ROWS_PER_PASS)
                if out is None:
                    out = torch.empty((B_all, self.out_elems), dtype=self.torch_dtype, device=self.device)
                for lo in range(0, B_all, tile):
                    hi = min(B_all, lo + tile)
                    self.run(None if ev_states is None else ev_states[lo:hi], out=out[lo:hi],
                             soft=None if soft is None else soft[lo:hi])
                return out
        if ev_states is None:
            if out is None and soft is None:
                raise ValueError("batch size unknown: pass ev_states or out")
            B = out.shape[0] if out is not None else soft.shape[0]
        else:
            if ev_states.dtype != torch.int32 or not ev_states.is_cuda or not ev_states.is_contiguous():
                raise ValueError("ev_states must be a contiguous int32 CUDA tensor")
            if ev_states.dim() != 2 or ev_states.shape[1] != self.n_ev:
                raise ValueError(f"ev_states must be [B, {self.n_ev}]")
            B = ev_states.shape[0]
        if B <= 0:
            raise ValueError("empty batch")
        self._sets_seen += B
        if not self._spec_tried and self.AUTO_SPECIALIZE_SETS is not None and self._sets_seen >= self.AUTO_SPECIALIZE_SETS \
                and soft is None and not getattr(self, "trace_cols", 0):
            try:
                self.specialize()
            except N.PgxError:
                pass  # divide / max steps, too large, no libnvrtc: the table-driven and step kernels keep serving it
        # (the device guard costs ~5 us of host time per call: skipped when the plan's device is already current — small
        # buckets of a mixed-evidence batch are bound by this host path, not by their kernels)
        guard = contextlib.nullcontext() if torch.cuda.current_device() == self.device.index else torch.cuda.device(self.device)
        with guard:
            if out is None:
                out = torch.empty((B, self.out_elems), dtype=self.torch_dtype, device=self.device)
            elif out.dtype != self.torch_dtype or not out.is_contiguous() or tuple(out.shape) != (B, self.out_elems):
                raise ValueError("out has the wrong dtype/shape")
            need = self.workspace_bytes(B)
            if workspace is None:
                if self._ws is None or self._ws.numel() < need:
                    self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
                workspace = self._ws
            elif workspace.numel() * workspace.element_size() < need:
                raise ValueError(f"workspace too small: need {need} bytes")
            stream = torch.cuda.current_stream(self.device).cuda_stream
            N.check(
                self.lib.pgx_run_batch_soft(
                    self.handle,
                    C.c_void_p(ev_states.data_ptr() if (ev_states is not None and self.n_ev) else 0),
                    C.c_void_p(soft.data_ptr() if soft is not None else 0),
                    C.c_void_p(out.data_ptr()),
                    C.c_void_p(workspace.data_ptr()),
                    workspace.numel() * workspace.element_size(),
                    B,
                    C.c_void_p(stream),
                )
            )
        return out

    def set_trace(self, trace: np.ndarray):
        """Attach the traceback descriptor of a max-product plan (planner.compile_jt_mpe_plan)."""
        trace = np.ascontiguousarray(trace, dtype=np.int32)
        N.check(self.lib.pgx_plan_set_trace(self.handle, trace.ctypes.data_as(C.POINTER(C.c_int32)), trace.size))
        self.trace_cols = int(trace[1])

    def run_mpe(self, ev_states, soft=None):
        """Max-product pass + traceback: int32 CUDA tensor [B, n_columns] of state indices (most probable explanation of
        every unobserved variable, column order of the descriptor). Batches whose workspace would exceed the cap are
        processed in row tiles."""
        torch = _torch()
        if not getattr(self, "trace_cols", 0):
            raise ValueError("not a max-product plan: call set_trace first")
        if self.n_ev:
            if ev_states.dtype != torch.int32 or not ev_states.is_cuda or not ev_states.is_contiguous() or \
                    ev_states.dim() != 2 or ev_states.shape[1] != self.n_ev:
                raise ValueError(f"ev_states must be a contiguous int32 CUDA tensor [B, {self.n_ev}]")
            B_all = int(ev_states.shape[0])
        else:
            B_all = int(soft.shape[0]) if soft is not None else 1
        if self.plan.in_elems and (soft is None or soft.shape != (B_all, self.plan.in_elems) or soft.dtype != self.torch_dtype):
            raise ValueError(f"soft must be a {self.dtype_name} CUDA tensor [B, {self.plan.in_elems}]")
        with torch.cuda.device(self.device):
            out = torch.empty((B_all, self.trace_cols), dtype=torch.int32, device=self.device)
            tile = B_all
            if B_all > 32 and self.workspace_bytes(B_all) > self.MAX_WORKSPACE_BYTES:
                per_set = self.workspace_bytes(32) // 32
                tile = max(32, int(self.MAX_WORKSPACE_BYTES // max(per_set, 1)) // 32 * 32)
            tile = min(tile, self.MAX_ROWS_PER_PASS)
            need = self.workspace_bytes(min(tile, B_all))
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            stream = torch.cuda.current_stream(self.device).cuda_stream
            for lo in range(0, B_all, tile):
                hi = min(B_all, lo + tile)
                ev = ev_states[lo:hi] if self.n_ev else None
                sf = soft[lo:hi].contiguous() if soft is not None else None
                N.check(self.lib.pgx_run_batch_mpe(
                    self.handle, C.c_void_p(ev.data_ptr() if ev is not None else 0), C.c_void_p(sf.data_ptr() if sf is not None else 0),
                    C.c_void_p(out[lo:hi].data_ptr()), C.c_void_p(self._ws.data_ptr()), self._ws.numel(), hi - lo, C.c_void_p(stream)))
        return out

    def profile_steps(self, ev_states):
        """Per-step device time (ms) of one stepwise pass: list of (ms, out_size, sum_size, n_operands, alg_bytes)."""
        torch = _torch()
        B = ev_states.shape[0]
        out = torch.empty((B, self.out_elems), dtype=self.torch_dtype, device=self.device)
        need = self.workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        n = self.plan.n_steps
        ms = (C.c_float * n)()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.pgx_profile_steps(self.handle, C.c_void_p(ev_states.data_ptr() if self.n_ev else 0),
                                           C.c_void_p(out.data_ptr()), C.c_void_p(self._ws.data_ptr()), self._ws.numel(), B,
                                           C.c_void_p(stream), ms, n))
        item = 8 if self.dtype_name == "float64" else 4
        rows = []
        for t, st in zip(ms, self.plan.steps):
            ssz = 1
            for v in st.sum_vars:
                ssz *= self.plan.card[v]
            work = st.out.size + sum(tb.size for tb, _ in st.operands if tb.kind == 1)
            const = sum(tb.size for tb, _ in st.operands if tb.kind == 0)
            rows.append((float(t), st.out.size, ssz, len(st.operands), item * (B * work + const), st.level))
        return rows

    def profile_launches(self, ev_states):
        """Device time of every launch of one stepwise pass with the production schedule (steps of a dependency level
        share a launch). Returns a list of dicts: ms, steps (indices into plan.steps), alg_bytes."""
        torch = _torch()
        B = ev_states.shape[0]
        out = torch.empty((B, self.out_elems), dtype=self.torch_dtype, device=self.device)
        need = self.workspace_bytes(B)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        n = max(1, self.plan.n_steps)
        ms = (C.c_float * n)()
        owner = (C.c_int32 * n)()
        n_l = C.c_int32(0)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        N.check(self.lib.pgx_profile_launches(self.handle, C.c_void_p(ev_states.data_ptr() if self.n_ev else 0),
                                              C.c_void_p(out.data_ptr()), C.c_void_p(self._ws.data_ptr()), self._ws.numel(), B,
                                              C.c_void_p(stream), ms, n, owner, n, C.byref(n_l)))
        item = 8 if self.dtype_name == "float64" else 4
        rows = [{"ms": float(ms[i]), "steps": [], "alg_bytes": 0} for i in range(n_l.value)]
        for si, st in enumerate(self.plan.steps):
            work = st.out.size + sum(tb.size for tb, _ in st.operands if tb.kind == 1)
            const = sum(tb.size for tb, _ in st.operands if tb.kind == 0)
            r = rows[owner[si]]
            r["steps"].append(si)
            r["alg_bytes"] += item * (B * work + const)
        return rows

    def run_pinned(self, ev_pinned, out_pinned, n_chunks: int = 0, n_streams: int = 3):
        """End-to-end call with HOST buffers: pinned int32 [B, n_ev] in, pinned [B, out_elems] out.

        The batch is cut into chunks that flow through a small ring of CUDA streams, so the H2D copy of chunk
        i+1, the kernels of chunk i and the D2H copy of chunk i-1 overlap (the posterior rows are ~36x larger
        than the evidence rows, so the D2H leg is what bounds the end-to-end rate over PCIe). Blocks until every
        chunk has landed in `out_pinned`."""
        torch = _torch()
        B = int(ev_pinned.shape[0]) if self.n_ev else int(out_pinned.shape[0])
        if n_chunks <= 0:
            # enough chunks that the first D2H copy starts early, few enough that the per-copy overhead stays small:
            # with the specialised kernel (whole batch in ~30 us) 4 chunks measured best (7.38e7 queries/s; 16: 7.18e7)
            n_chunks = max(1, min(4 if self._spec_tried and self.info(N.INFO_SPECIALIZED) else 16, B // 8192))
        n_streams = max(1, min(n_streams, n_chunks))
        chunk = -(-B // n_chunks)
        chunk = -(-chunk // 32) * 32
        with torch.cuda.device(self.device):
            ring = getattr(self, "_ring", None)
            if ring is None or ring["chunk"] < chunk or len(ring["streams"]) < n_streams:
                ring = {
                    "chunk": chunk,
                    "streams": [torch.cuda.Stream(self.device) for _ in range(n_streams)],
                    "ev": [torch.empty((chunk, max(self.n_ev, 1)), dtype=torch.int32, device=self.device) for _ in range(n_streams)],
                    "out": [torch.empty((chunk, self.out_elems), dtype=self.torch_dtype, device=self.device) for _ in range(n_streams)],
                    "ws": [torch.empty(self.workspace_bytes(chunk), dtype=torch.uint8, device=self.device) for _ in range(n_streams)],
                }
                self._ring = ring
            cur = torch.cuda.current_stream(self.device)
            for st in ring["streams"][:n_streams]:
                st.wait_stream(cur)
            for c in range(n_chunks):
                lo, hi = c * chunk, min(B, (c + 1) * chunk)
                if lo >= hi:
                    break
                k = c % n_streams
                with torch.cuda.stream(ring["streams"][k]):
                    ev_d = None
                    if self.n_ev:
                        ev_d = ring["ev"][k][: hi - lo, : self.n_ev]
                        if self.n_ev != ring["ev"][k].shape[1]:
                            ev_d = ring["ev"][k].view(-1)[: (hi - lo) * self.n_ev].view(hi - lo, self.n_ev)
                        ev_d.copy_(ev_pinned[lo:hi], non_blocking=True)
                    out_d = ring["out"][k][: hi - lo]
                    self.run(ev_d, out=out_d, workspace=ring["ws"][k])
                    out_pinned[lo:hi].copy_(out_d, non_blocking=True)
            for st in ring["streams"][:n_streams]:
                st.synchronize()
        return out_pinned

    def run_host(self, ev_states_np: np.ndarray, soft=None) -> np.ndarray:
        """Convenience: host int array [B, n_ev] (+ host soft-evidence rows [B, in_elems]) -> host posteriors
        [B, out_elems] (range-checks the states)."""
        torch = _torch()
        if soft is not None:
            soft = torch.from_numpy(np.ascontiguousarray(soft, dtype=np.float64)).to(self.device, dtype=self.torch_dtype)
        ev = np.ascontiguousarray(ev_states_np, dtype=np.int32).reshape(-1, max(self.n_ev, 1))
        cards = [self.plan.card[v] for v in self.plan.ev_vars]
        for j, c in enumerate(cards):
            col = ev[:, j]
            if col.size and (col.min() < 0 or col.max() >= c):
                raise ValueError(f"evidence state out of range for {self.plan.ev_vars[j]}")
        if self.n_ev == 0:
            B = max(1, int(np.asarray(ev_states_np).shape[0])) if np.asarray(ev_states_np).ndim >= 1 else 1
            out = torch.empty((B, self.out_elems), dtype=self.torch_dtype, device=self.device)
            return self.run(None, out=out, soft=soft).cpu().numpy()
        dev = torch.from_numpy(ev).to(self.device)
        return self.run(dev, soft=soft).cpu().numpy()
