"""Static contraction plans: the POD hand-off between the Python planner and the CUDA engine.

A plan is ONE int32 word pool (layout below, mirrored in include/pgx.h) plus ONE packed table blob of
batch-invariant values (CPTs / clique potentials, reference layout: C-order, first variable slowest,
pgmpy/factors/discrete/DiscreteFactor.py:91-127).

Every step is the same operation, a fused product + sum-out over a batch of evidence sets b:

    out[o, b] = REDUCE_s  PROD_k  operand_k[ idx_k(o, s) + evoff_k(b) , b ]   ( / divisor_j[idx_j(o), b] )

  * `o` runs over the output scope (row-major over `out dims`), `s` over the summed-out scope;
  * idx_k is a mixed-radix dot product with per-operand strides (stride 0 = variable not in scope):
    this replaces the reference's `einsum` product (DiscreteFactor.py:769-777) + `einsum` sum-out
    (:400-408) without materialising the joint table;
  * evoff_k(b) = sum_j ev_states[b, slot_j] * stride_j is the evidence reduce (:599-614) folded into
    operand addressing (the observed axes never appear in any step's scope);
  * divisors implement DiscreteFactor.divide (:838-863; 0/0 -> 0, x/0 -> inf) after the reduction;
  * work tables (batch dependent) live in the workspace as [entry][b] (b fastest), const tables in
    the blob; the last steps' outputs are copied/normalised into out[b, :] by segments
    (normalise = DiscreteFactor.normalize, :530).

Word pool layout (int32 words; 64-bit values as lo,hi):
  header[16]: 0 magic 'PGX1' | 1 version | 2 n_ev | 3 n_steps | 4 n_segs | 5 out_elems | 6,7 ws_entries
              | 8,9 const_entries | 10 step_index_off | 11 segs_off | 12 max_ops | 13 max_axes | 14 ev_card_off
              | 15 inputs_off (0: the plan has no batch-dependent input tables)
  ev_card[n_ev]: cardinality of each evidence slot (states are range-checked against it)
  step_index[n_steps]: word offset of each step record
  step record: 0 A (#out axes) | 1 S (#sum axes) | 2 K (#operands) | 3 flags (bit0 max-reduce, bit1 has divisor)
               | 4,5 out_size | 6,7 sum_size | 8,9 out work offset | 10 level | 11 0
               | out dims[A] | sum dims[S] | K operand records | evidence pairs
  operand record (6+A+S words): 0 kind (0 const, 1 work; bit8 divisor) | 1,2 table offset | 3 n_ev
               | 4 ev_off (words from step start) | 5 0 | strides over out axes[A] | strides over sum axes[S]
  evidence pair: (slot, stride)
  segment record[8]: 0,1 work offset | 2 size | 3 out offset | 4 flags (bit0 normalise) | 5,6,7 0
  inputs block (at inputs_off): n_inputs | in_elems | n_inputs x (0,1 work offset | 2 size | 3 offset in the input row)
      Batch-dependent INPUT tables: work tables the caller fills per evidence set (soft / virtual evidence: one
      likelihood vector per variable, the factor pgmpy/inference/base.py:256-299 adds as an observed binary child).
      The engine copies row b of the caller's `soft[B, in_elems]` into them before the first step.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

MAGIC = 0x50475831
VERSION = 1
HEADER_WORDS = 16
STEP_FIXED = 12
OP_FIXED = 6
SEG_WORDS = 8
MAX_OPS = 16  # operands per step the kernels are compiled for; larger products are split by the builder
MAX_AXES = 24  # out axes + sum axes per step after coalescing

KIND_CONST = 0
KIND_WORK = 1
FLAG_MAX = 1
FLAG_DIV = 2
SEG_NORMALIZE = 1


def _lohi(x: int) -> Tuple[int, int]:
    x = int(x)
    lo = x & 0xFFFFFFFF
    hi = (x >> 32) & 0xFFFFFFFF
    return (lo - (1 << 32) if lo >= (1 << 31) else lo, hi - (1 << 32) if hi >= (1 << 31) else hi)


def lohi_to_int(lo: int, hi: int) -> int:
    return (int(hi) << 32) | (int(lo) & 0xFFFFFFFF)


@dataclass
class Table:
    """A dense table over `vars` (C-order). kind const -> offset into the blob; work -> workspace."""

    kind: int
    vars: Tuple[Hashable, ...]
    dims: Tuple[int, ...]
    offset: int = -1  # entries; work tables get theirs in finalize()
    tid: int = -1
    first_step: int = -1
    last_step: int = -1
    is_input: bool = False  # work table filled from the caller's per-evidence-set input row before the first step

    @property
    def size(self) -> int:
        s = 1
        for d in self.dims:
            s *= int(d)
        return s


@dataclass
class StepSpec:
    out: Table
    sum_vars: Tuple[Hashable, ...]
    operands: List[Tuple[Table, bool]]  # (table, is_divisor)
    reduce_max: bool = False
    level: int = 0


@dataclass
class Segment:
    table: Table
    out_offset: int
    normalize: bool
    vars: Tuple[Hashable, ...] = ()


@dataclass
class Plan:
    pool: np.ndarray  # int32
    const_blob: np.ndarray  # float64
    ev_vars: Tuple[Hashable, ...]
    segments: List[Segment]
    out_elems: int
    ws_entries: int
    n_steps: int
    card: Dict[Hashable, int]
    steps: List[StepSpec] = field(default_factory=list, repr=False)
    meta: dict = field(default_factory=dict)
    inputs: List[Tuple[Tuple[Hashable, ...], int, int]] = field(default_factory=list)  # (vars, size, offset in the input row)
    in_elems: int = 0

    def algorithmic_bytes(self, batch: int, itemsize: int = 8) -> int:
        """every tensor counted once per step in which it is an operand or result,
        batch-invariant operands once per step, plus evidence in and posteriors out."""
        if not self.steps and self.meta.get("alg_bytes_per_set"):
            return int(self.meta["alg_bytes_per_set"]) * batch  # plan loaded from the cache: figure stored per set
        total = 0
        for st in self.steps:
            work = st.out.size
            const = 0
            for t, _ in st.operands:
                if t.kind == KIND_WORK:
                    work += t.size
                else:
                    const += t.size
            total += itemsize * (batch * work + const)
        total += 4 * batch * len(self.ev_vars) + itemsize * batch * (self.out_elems + 2 * self.in_elems)
        return total

    # ---- plan cache on disk: a compiled plan is two arrays + a little metadata -----------
    def save(self, path: str) -> None:
        """Serialise the executable part of the plan (word pool, table blob, evidence slots, output segments).
        A loaded plan runs exactly like the original; only the planner-side step objects are not kept."""
        import json

        segs = [[list(map(str, s.vars)), int(s.out_offset), int(s.table.size), bool(s.normalize), list(map(int, s.table.dims))]
                for s in self.segments]
        header = {
            "ev_vars": list(self.ev_vars), "card": {str(k): int(v) for k, v in self.card.items()},
            "segments": segs, "out_elems": self.out_elems, "ws_entries": self.ws_entries, "n_steps": self.n_steps,
            "meta": {k: v for k, v in self.meta.items() if isinstance(v, (str, int, float, bool))},
            "alg_bytes_per_set": self.algorithmic_bytes(1) if self.steps else self.meta.get("alg_bytes_per_set", 0),
            "inputs": [[list(map(str, v)), int(n), int(o)] for v, n, o in self.inputs], "in_elems": int(self.in_elems),
        }
        np.savez_compressed(path, pool=self.pool, const_blob=self.const_blob, header=np.array(json.dumps(header)))

    @classmethod
    def load(cls, path: str) -> "Plan":
        import json

        with np.load(path, allow_pickle=False) as z:
            header = json.loads(str(z["header"]))
            pool = z["pool"].astype(np.int32)
            blob = z["const_blob"].astype(np.float64)
        segments = []
        for vars_, out_off, size, norm, dims in header["segments"]:
            t = Table(KIND_WORK, tuple(vars_), tuple(dims))
            segments.append(Segment(t, out_off, norm, tuple(vars_)))
        meta = dict(header["meta"])
        meta["alg_bytes_per_set"] = header["alg_bytes_per_set"]
        return cls(pool=pool, const_blob=blob, ev_vars=tuple(header["ev_vars"]), segments=segments,
                   out_elems=header["out_elems"], ws_entries=header["ws_entries"], n_steps=header["n_steps"],
                   card=header["card"], steps=[], meta=meta,
                   inputs=[(tuple(v), n, o) for v, n, o in header.get("inputs", [])], in_elems=header.get("in_elems", 0))

    def operand_loads(self) -> int:
        """Operand loads (= multiplies) per evidence set: sum over steps of |out| * |sum| * #operands."""
        total = 0
        for st in self.steps:
            joint = st.out.size
            for v in st.sum_vars:
                joint *= self.card[v]
            total += joint * max(1, len(st.operands))
        return total

    def flops(self, batch: int) -> int:
        f = 0
        for st in self.steps:
            joint = st.out.size
            for v in st.sum_vars:
                joint *= self.card[v]
            f += joint * max(1, len(st.operands))
        return f * batch


class PlanBuilder:
    """Collects tables and steps, then assigns workspace offsets by liveness and packs the pool."""

    def __init__(self, card: Dict[Hashable, int], ev_vars: Sequence[Hashable], reassociate: bool = True):
        self.reassociate = reassociate and os.environ.get("PGX_NO_REASSOC", "0") != "1"
        self.card = {v: int(c) for v, c in card.items()}
        self.ev_vars = tuple(ev_vars)
        self.ev_slot = {v: i for i, v in enumerate(self.ev_vars)}
        self.tables: List[Table] = []
        self.steps: List[StepSpec] = []
        self.segments: List[Segment] = []
        self._const_chunks: List[np.ndarray] = []
        self._const_len = 0
        self._const_cache: Dict[int, Table] = {}
        self._memo: Dict[tuple, Table] = {}  # contraction -> its result table (common-subexpression reuse)
        self.inputs: List[Table] = []

    # ---- tables ----------------------------------------------------------------------------
    def add_const(self, vars_: Sequence[Hashable], values: np.ndarray, key=None) -> Table:
        if key is not None and key in self._const_cache:
            return self._const_cache[key]
        vars_ = tuple(vars_)
        dims = tuple(self.card[v] for v in vars_)
        values = np.ascontiguousarray(np.asarray(values, dtype=np.float64))
        if values.size != int(np.prod(dims, dtype=np.int64)):
            raise ValueError("const table size does not match its scope")
        t = Table(KIND_CONST, vars_, dims, offset=self._const_len, tid=len(self.tables))
        self._const_chunks.append(values.reshape(-1))
        # keep every table 16-byte aligned for both dtypes
        pad = (-values.size) % 4
        if pad:
            self._const_chunks.append(np.zeros(pad))
        self._const_len += values.size + pad
        self.tables.append(t)
        if key is not None:
            self._const_cache[key] = t
        return t

    def new_work(self, vars_: Sequence[Hashable]) -> Table:
        vars_ = tuple(vars_)
        for v in vars_:
            if v in self.ev_slot:
                raise ValueError(f"work table scope may not contain evidence variable {v}")
        t = Table(KIND_WORK, vars_, tuple(self.card[v] for v in vars_), tid=len(self.tables))
        self.tables.append(t)
        return t

    def add_input(self, vars_: Sequence[Hashable]) -> Table:
        """A batch-dependent table the CALLER supplies per evidence set (soft evidence likelihoods). It is a work
        table that is alive from before the first step; its values arrive as a slice of the input row."""
        t = self.new_work(vars_)
        t.is_input = True
        self.inputs.append(t)
        return t

    # ---- steps -----------------------------------------------------------------------------
    # cost-model knobs of contract()
    SPLIT_MIN_JOINT = 1 << 15  # steps smaller than this are never split
    SPLIT_MIN_OUT = 1024       # ... and a step should expose at least this many output entries to the grid

    def _free(self, t: Table) -> List[Hashable]:
        return [v for v in t.vars if v not in self.ev_slot]

    def _prod(self, vars_) -> int:
        s = 1
        for v in vars_:
            s *= self.card[v]
        return s

    def contract(
        self,
        operands: Sequence[Table],
        out_vars: Sequence[Hashable],
        divisors: Sequence[Table] = (),
        reduce_max: bool = False,
        level: int = 0,
        optimize: bool = True,
        split: bool = True,
    ) -> Table:
        """out[out_vars] = reduce over every other non-evidence variable of prod(operands) / prod(divisors).

        With optimize=True the product is first re-associated greedily (like the pairwise path the reference
        gets from opt_einsum's "greedy", pgmpy/inference/ExactInference.py:404): two operands are multiplied in
        a cheap step of their own (summing variables nobody else needs) whenever that is cheaper than dragging
        one more operand through the final pass; and a final pass with few output entries but a huge summed
        range is cut in two (partial sums over an enlarged output, then a small reduce) so the grid has work."""
        out_vars = tuple(out_vars)
        ops = list(operands)
        # identical contractions are computed once (tables are written once and never change): the distribute pass
        # re-derives many of the partial products the collect pass already built
        memo_key = (tuple(sorted(t.tid for t in ops)), tuple(t.tid for t in divisors), out_vars, bool(reduce_max))
        hit = self._memo.get(memo_key)
        if hit is not None:
            return hit
        result = self._contract(ops, out_vars, divisors, reduce_max, level, optimize, split)
        self._memo[memo_key] = result
        return result

    def _contract(self, ops, out_vars, divisors, reduce_max, level, optimize, split) -> Table:
        needed_later = set(out_vars)
        for d in divisors:
            needed_later |= set(self._free(d))

        def union_scope(tables):
            sc: List[Hashable] = []
            for t in tables:
                for v in self._free(t):
                    if v not in sc:
                        sc.append(v)
            return sc

        if optimize and len(ops) > 2 and self.reassociate:
            # Greedy pairwise re-association by operand loads. Evaluating the product directly costs one load per
            # operand per entry of the joint scope J: J * n. Multiplying two operands first costs 2 loads per entry of
            # THEIR joint scope, a store and a re-load of what is kept, and leaves n - 1 operands over a joint scope that
            # has lost the variables summed away in the pair step.
            while len(ops) > 2:
                scope_all = union_scope(ops)
                final_joint = self._prod(scope_all)
                direct = final_joint * len(ops)
                count: Dict[Hashable, int] = {}
                for t in ops:
                    for v in self._free(t):
                        count[v] = count.get(v, 0) + 1
                best = None
                for i in range(len(ops)):
                    fi = self._free(ops[i])
                    si = set(fi)
                    for j in range(i + 1, len(ops)):
                        fj = self._free(ops[j])
                        sj = set(fj)
                        u = fi + [v for v in fj if v not in si]
                        cost = self._prod(u)
                        if 2 * cost >= direct:
                            continue
                        keep = [v for v in u if v in needed_later or count[v] > (1 if v in si else 0) + (1 if v in sj else 0)]
                        kept = self._prod(keep)
                        rest_joint = final_joint // (cost // kept)
                        total = 2 * cost + 2 * kept + rest_joint * (len(ops) - 1)
                        key = (total, kept, cost)
                        if best is None or key < best[0]:
                            best = (key, i, j, keep)
                if best is None or best[0][0] >= direct:
                    break
                _, i, j, keep = best
                merged = self.contract([ops[i], ops[j]], keep, reduce_max=reduce_max, level=level, optimize=False)
                ops = [t for k, t in enumerate(ops) if k not in (i, j)] + [merged]
        # hard limit of the kernels: no step exceeds MAX_OPS operands (divisors count)
        while len(ops) + len(divisors) > MAX_OPS:
            ops.sort(key=lambda t: t.size)
            group, ops = ops[:MAX_OPS], ops[MAX_OPS:]
            later = set(needed_later)
            for t in ops:
                later |= set(t.vars)
            keep = [v for v in union_scope(group) if v in later]
            ops.append(self.contract(group, keep, reduce_max=reduce_max, level=level, optimize=False))
        scope = union_scope(ops)
        for v in out_vars:
            if v not in scope:
                raise ValueError(f"output variable {v} is not in any operand")
        for d in divisors:
            for v in self._free(d):
                if v not in out_vars:
                    raise ValueError("divisor scope must be within the output scope")
        sum_vars = tuple(v for v in scope if v not in out_vars)
        if split and sum_vars:
            out_size = self._prod(out_vars)
            joint = out_size * self._prod(sum_vars)
            if joint >= self.SPLIT_MIN_JOINT and out_size < self.SPLIT_MIN_OUT:
                # keep some summed variables as extra output axes of a partial step
                keep: List[Hashable] = []
                size = out_size
                for v in sorted(sum_vars, key=lambda v: -self.card[v]):
                    if size >= self.SPLIT_MIN_OUT:
                        break
                    keep.append(v)
                    size *= self.card[v]
                if keep and len(keep) < len(sum_vars):
                    partial = self.contract(ops, out_vars + tuple(keep), reduce_max=reduce_max, level=level, optimize=False,
                                            split=False)
                    return self.contract([partial], out_vars, divisors=divisors, reduce_max=reduce_max, level=level,
                                         optimize=False, split=False)
        ops.sort(key=lambda t: t.kind)  # batch-invariant operands first (the fused kernel specialises on that)
        out = self.new_work(out_vars)
        idx = len(self.steps)
        self.steps.append(
            StepSpec(out, sum_vars, [(t, False) for t in ops] + [(t, True) for t in divisors], reduce_max, level)
        )
        out.first_step = idx
        for t in list(ops) + list(divisors):
            if t.kind == KIND_WORK:
                t.last_step = max(t.last_step, idx)
        return out

    # ---- trial builds ------------------------------------------------------------------------
    # cost model of a step sequence, per evidence set: HBM time of its work-table traffic + issue time of its operand
    # loads (rates measured on B200 for the step kernels: ~4.5 TB/s streaming, ~3e12 operand loads/s)
    COST_BYTES_PER_NS = 4500.0
    COST_LOADS_PER_NS = 3000.0

    def mark(self):
        """Snapshot for rollback(): lets the planner build a candidate step sequence, price it, and undo it."""
        return (len(self.tables), len(self.steps), len(self.segments), len(self._const_chunks), self._const_len,
                dict(self._const_cache), dict(self._memo), [t.last_step for t in self.tables], len(self.inputs))

    def rollback(self, mk) -> None:
        n_t, n_s, n_g, n_c, c_len, c_cache, memo, last, n_in = mk
        del self.tables[n_t:], self.steps[n_s:], self.segments[n_g:], self._const_chunks[n_c:], self.inputs[n_in:]
        self._const_len = c_len
        self._const_cache = dict(c_cache)  # copies: the same mark may be rolled back to more than once
        self._memo = dict(memo)
        for t, l in zip(self.tables, last):
            t.last_step = l

    def cost_since(self, mk) -> float:
        """Estimated ns per evidence set of the steps added since mark()."""
        loads = 0
        nbytes = 0
        for st in self.steps[mk[1]:]:
            joint = st.out.size * self._prod(st.sum_vars)
            loads += joint * max(1, len(st.operands))
            nbytes += 8 * (st.out.size + sum(t.size for t, _ in st.operands if t.kind == KIND_WORK))
        return nbytes / self.COST_BYTES_PER_NS + loads / self.COST_LOADS_PER_NS

    def emit(self, table: Table, normalize: bool, vars_: Sequence[Hashable] = None) -> Segment:
        if table.kind != KIND_WORK:
            raise ValueError("only work tables can be emitted")
        off = sum(s.table.size for s in self.segments)
        seg = Segment(table, off, normalize, tuple(vars_ if vars_ is not None else table.vars))
        self.segments.append(seg)
        table.last_step = 1 << 60  # alive until the output pass
        return seg

    # ---- lowering --------------------------------------------------------------------------
    def _strides(self, t: Table) -> Dict[Hashable, int]:
        st = {}
        acc = 1
        for v, d in zip(reversed(t.vars), reversed(t.dims)):
            st[v] = acc
            acc *= d
        return st

    def _lower_step(self, st: StepSpec) -> List[int]:
        out_vars = [v for v in st.out.vars]
        out_dims = [self.card[v] for v in out_vars]
        sum_vars = list(st.sum_vars)
        sum_dims = [self.card[v] for v in sum_vars]
        per_op = []
        for t, is_div in st.operands:
            strides = self._strides(t)
            so = [strides.get(v, 0) for v in out_vars]
            ss = [strides.get(v, 0) for v in sum_vars]
            ev = [(self.ev_slot[v], strides[v]) for v in t.vars if v in self.ev_slot]
            for v in t.vars:
                if v not in self.ev_slot and v not in out_vars and v not in sum_vars:
                    raise AssertionError("operand variable not covered by the step")
            per_op.append((t, is_div, so, ss, ev))

        def coalesce(dims, cols):
            """merge adjacent axes (i, i+1) when every operand walks them contiguously; drop extent-1 axes"""
            keep = [i for i, d in enumerate(dims) if d != 1]
            dims = [dims[i] for i in keep]
            cols = [[c[i] for i in keep] for c in cols]
            i = 0
            while i + 1 < len(dims):
                if all(c[i] == c[i + 1] * dims[i + 1] for c in cols):
                    dims[i] = dims[i] * dims[i + 1]
                    del dims[i + 1]
                    for c in cols:
                        c[i] = c[i + 1]
                        del c[i + 1]
                else:
                    i += 1
            return dims, cols

        # the output itself is an (implicit) operand of the out axes: row-major contiguous
        out_self = []
        acc = 1
        for d in reversed(out_dims):
            out_self.insert(0, acc)
            acc *= d
        odims, ocols = coalesce(out_dims, [p[2] for p in per_op] + [out_self])
        ocols = ocols[:-1]
        sdims, scols = coalesce(sum_dims, [p[3] for p in per_op])
        A, S, K = len(odims), len(sdims), len(per_op)
        if A + S > MAX_AXES:
            raise ValueError(f"step needs {A + S} axes > MAX_AXES={MAX_AXES}")
        out_size = st.out.size
        sum_size = 1
        for d in sum_dims:
            sum_size *= d
        flags = (FLAG_MAX if st.reduce_max else 0) | (FLAG_DIV if any(p[1] for p in per_op) else 0)
        rec = [A, S, K, flags, *_lohi(out_size), *_lohi(sum_size), *_lohi(st.out.offset), st.level, 0]
        rec += odims + sdims
        op_words = OP_FIXED + A + S
        ev_base = STEP_FIXED + A + S + K * op_words
        ev_words: List[int] = []
        for k, (t, is_div, _, _, ev) in enumerate(per_op):
            kind = t.kind | (0x100 if is_div else 0)
            rec += [kind, *_lohi(t.offset), len(ev), ev_base + len(ev_words), 0]
            rec += ocols[k] + scols[k]
            for slot, stride in ev:
                ev_words += [slot, stride]
        rec += ev_words
        return rec

    def finalize(self, meta: Optional[dict] = None) -> Plan:
        # 1. dependency levels (longest path): a step's level is one more than the deepest producer of its
        #    work operands. Steps of one level are mutually independent, so the fused kernel may run them
        #    concurrently; sorting by level keeps the sequential order valid for the stepwise path.
        producer = {st.out.tid: i for i, st in enumerate(self.steps)}
        level = [0] * len(self.steps)
        for i, st in enumerate(self.steps):
            lv = 0
            for t, _ in st.operands:
                if t.kind == KIND_WORK and not t.is_input:
                    lv = max(lv, level[producer[t.tid]] + 1)
            level[i] = lv
            st.level = lv
        order = sorted(range(len(self.steps)), key=lambda i: (level[i], i))
        self.steps = [self.steps[i] for i in order]
        n_steps = len(self.steps)
        # 2. liveness in units of levels: a table is born at its producer's level and dies after the last
        #    level that reads it (emitted tables never die)
        for t in self.tables:
            t.first_step = 0 if t.is_input else -1  # inputs are written before level 0 runs
            if t.last_step < (1 << 59):
                t.last_step = -1
        for st in self.steps:
            st.out.first_step = st.level
            for t, _ in st.operands:
                if t.kind == KIND_WORK and t.last_step < (1 << 59):
                    t.last_step = max(t.last_step, st.level)
        work = [t for t in self.tables if t.kind == KIND_WORK and t.first_step >= 0]
        for t in work:
            if t.last_step < 0:
                t.last_step = t.first_step
        by_birth = sorted(work, key=lambda t: (t.first_step, not t.is_input, t.tid))
        free: List[Tuple[int, int]] = []  # (offset, size), kept sorted by offset
        live: List[Table] = []
        top = 0

        def release(t):
            nonlocal free
            free.append((t.offset, t.size))
            free.sort()
            merged = []
            for off, sz in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1] = (merged[-1][0], merged[-1][1] + sz)
                else:
                    merged.append((off, sz))
            free = merged

        for t in by_birth:
            # memory is recycled only from tables whose last reader ran in a strictly earlier level
            still = []
            for l in live:
                if l.last_step < t.first_step:
                    release(l)
                else:
                    still.append(l)
            live = still
            need = max(1, t.size)
            placed = False
            best = None
            for i, (off, sz) in enumerate(free):  # best fit
                if sz >= need and (best is None or sz < free[best][1]):
                    best = i
            if best is not None:
                off, sz = free[best]
                t.offset = off
                if sz == need:
                    free.pop(best)
                else:
                    free[best] = (off + need, sz - need)
                placed = True
            if not placed:
                if free and free[-1][0] + free[-1][1] == top:
                    off, sz = free.pop()
                    t.offset = off
                    top = off + need
                else:
                    t.offset = top
                    top += need
            live.append(t)
        ws_entries = max(1, top)

        step_recs = [self._lower_step(st) for st in self.steps]
        out_elems = sum(s.table.size for s in self.segments)
        max_ops = max([len(st.operands) for st in self.steps], default=0)
        max_axes = max([r[0] + r[1] for r in step_recs], default=0)
        const_len = max(self._const_len, 1)
        header = [MAGIC, VERSION, len(self.ev_vars), n_steps, len(self.segments), out_elems]
        header += [*_lohi(ws_entries), *_lohi(const_len)]
        ev_card_off = HEADER_WORDS
        step_index_off = ev_card_off + len(self.ev_vars)
        pos = step_index_off + n_steps
        index = []
        for r in step_recs:
            index.append(pos)
            pos += len(r)
        segs_off = pos
        inputs_off = segs_off + SEG_WORDS * len(self.segments) if self.inputs else 0
        header += [step_index_off, segs_off, max_ops, max_axes, ev_card_off, inputs_off]
        words = header + [self.card[v] for v in self.ev_vars] + index
        for r in step_recs:
            words += r
        for s in self.segments:
            words += [*_lohi(s.table.offset), s.table.size, s.out_offset, SEG_NORMALIZE if s.normalize else 0, 0, 0, 0]
        in_elems = 0
        plan_inputs = []
        if self.inputs:
            words += [len(self.inputs), sum(t.size for t in self.inputs)]
            for t in self.inputs:
                words += [*_lohi(t.offset), t.size, in_elems]
                plan_inputs.append((t.vars, t.size, in_elems))
                in_elems += t.size
        pool = np.array(words, dtype=np.int64)
        if pool.max(initial=0) >= (1 << 31) or pool.min(initial=0) < -(1 << 31):
            raise ValueError("plan word does not fit int32")
        blob = np.concatenate(self._const_chunks) if self._const_chunks else np.zeros(1)
        return Plan(
            pool=pool.astype(np.int32),
            const_blob=np.ascontiguousarray(blob, dtype=np.float64),
            ev_vars=self.ev_vars,
            segments=list(self.segments),
            out_elems=out_elems,
            ws_entries=ws_entries,
            n_steps=n_steps,
            card=dict(self.card),
            steps=list(self.steps),
            meta=dict(meta or {}),
            inputs=plan_inputs,
            in_elems=in_elems,
        )
