"""`VariableElimination` and `BeliefPropagation` with the reference's signatures, executed on the GPU.

Drop-in boundary: same constructor / query / calibrate signatures, same result types
(DiscreteFactor with `variables` in the caller's order and state names from the model), same error
conventions as pgmpy/inference/ExactInference.py:246-457 (VE) and :739-1220 (BP). Every numeric
result comes from libpgx.so; the per-query Python of the reference (model copies, re-validation,
pruning walks) is replaced by a plan cache keyed on the query signature.

Batched extension (ours): `query_batch(variables, evidence_vars, evidence_states)` evaluates one
signature for B evidence sets at once and returns a CUDA tensor [B, prod(card(variables))] (or, for
`BeliefPropagation.marginals_batch`, the concatenated marginals of every unobserved variable).
"""
from __future__ import annotations

import warnings
from typing import Dict, Hashable, List, Optional, Sequence

import numpy as np

from .config import config, default_dtype, logger, normalize_dtype
from .engine import CompiledPlan, require_cuda
from .factors import DiscreteFactor, TabularCPD
from .models import DiscreteBayesianNetwork, JunctionTree, from_pgmpy, junction_tree_from_pgmpy
from . import planner as PL


def _soft_items(inf, virtual_evidence):
    """{variable: (state names or None, likelihoods [B or 1, card])} with the reference's argument checks
    (pgmpy/inference/base.py:214-254)."""
    items = {}
    for item in virtual_evidence:
        if isinstance(item, tuple) and len(item) == 2 and not hasattr(item, "variables"):
            var, vals = item
            vals = np.asarray(vals.detach().cpu().numpy() if hasattr(vals, "detach") else vals, dtype=np.float64)
            vals = vals.reshape(1, -1) if vals.ndim == 1 else vals
            names = None
        else:
            if not isinstance(item, (TabularCPD, DiscreteFactor)) and not hasattr(item, "variables"):
                raise ValueError(
                    f"Virtual evidence should be an instance of TabularCPD or DiscreteFactor. Got: {type(item)}"
                )
            if len(item.variables) > 1:
                raise ValueError("Virtual evidence should be defined on individual variables.")
            var = item.variables[0]
            vals = np.asarray(item.values, dtype=np.float64).reshape(1, -1)
            names = list(item.state_names[var]) if getattr(item, "state_names", None) else None
        if var not in inf.variables:
            raise ValueError("Evidence provided for variable which is not in the model")
        if vals.ndim != 2 or vals.shape[1] != inf.cardinality[var]:
            raise ValueError(
                "The number of states/cardinality for the evidence should"
                " be same as the number of states/cardinality of the variable in the model"
            )
        if names is not None and list(names) != list(inf.states[var]):
            # the reference's child CPD carries the evidence's own state names for the parent; align by name
            order = [names.index(s) for s in inf.states[var]]
            vals = vals[:, order]
        items[var] = (names, vals)
    return items


def _as_bn(model):
    if isinstance(model, (DiscreteBayesianNetwork, JunctionTree)):
        return model
    name = type(model).__name__
    if name in ("LinearGaussianBayesianNetwork", "FunctionalBayesianNetwork"):
        return model  # rejected in query() with NotImplementedError like the reference
    if name == "JunctionTree":
        return junction_tree_from_pgmpy(model)
    if hasattr(model, "get_cpds") and hasattr(model, "edges"):
        return from_pgmpy(model)
    raise TypeError(f"unsupported model type {type(model)}")


class _Inference:
    def __init__(self, model, dtype: Optional[str] = None):
        # dtype None: pgmpy_b200.config, then the application's pgmpy.config when pgmpy is imported, then float64
        dtype = normalize_dtype(dtype) if dtype is not None else default_dtype()
        self._orig_model = model
        self.model = _as_bn(model)
        if type(self.model).__name__ in ("LinearGaussianBayesianNetwork", "FunctionalBayesianNetwork"):
            self._unsupported = type(self.model).__name__
            return
        self._unsupported = None
        self.model.check_model()  # inference/base.py:79-86
        self.dtype = dtype
        self._plans: Dict[tuple, CompiledPlan] = {}
        if isinstance(self.model, JunctionTree):
            self.variables = set(v for c in self.model.nodes() for v in c)
        else:
            self.variables = set(self.model.nodes())
        self.cardinality = self.model.get_cardinality()
        self.states = self.model.states

    def _compile(self, plan) -> CompiledPlan:
        """Upload a plan to the device `config` names (default: the current CUDA device), in this object's dtype."""
        logger.debug("pgmpy_b200: plan %s, %d steps, %d work entries, dtype %s", plan.meta.get("mode"), plan.n_steps,
                     plan.ws_entries, self.dtype)
        return CompiledPlan(plan, self.dtype, device=config.device_index())

    # ---- shared checks ---------------------------------------------------------------------
    def _check_query(self, variables, evidence, allow_empty=False):
        if self._unsupported:
            raise NotImplementedError(
                f"Variable Elimination is not supported for {self._unsupported}."
                f"Please use the 'predict' method of the {self._unsupported} class instead."
            )
        if isinstance(variables, str):
            raise TypeError("variables must be a list of strings")
        if isinstance(evidence, str):
            raise TypeError("evidence must be a list of strings")
        evidence = evidence if evidence is not None else {}
        common = set(evidence).intersection(set(variables))
        if common:
            raise ValueError(
                f"Can't have the same variables in both `variables` and `evidence`. Found in both: {common}"
            )
        if not variables and not allow_empty:
            raise ValueError("The `variables` argument to query() must contain at least one variable.")
        for v in list(variables) + list(evidence):
            if v not in self.variables:
                raise ValueError(f"Node {v} not in graph")
        return dict(evidence)

    def _states_of(self, ev_vars, evidence_rows) -> np.ndarray:
        return PL.evidence_to_states(self.states, ev_vars, evidence_rows)

    def _to_factor(self, variables, values) -> DiscreteFactor:
        card = [self.cardinality[v] for v in variables]
        return DiscreteFactor(list(variables), card, values, {v: self.states[v] for v in variables})

    def _soft_rows(self, cp: CompiledPlan, virtual_evidence, batch=None):
        """Soft (virtual) evidence -> the plan's input rows [B, in_elems] on the device. `virtual_evidence`: a list of
        TabularCPD / DiscreteFactor over one variable each (the reference's type, inference/base.py:256-299: one
        likelihood vector, B = 1) or of (variable, array [B, card]) pairs (batched: one vector per evidence set)."""
        torch = require_cuda()
        items = _soft_items(self, virtual_evidence)
        B = batch
        for _, vals in items.values():
            B = vals.shape[0] if B is None or vals.shape[0] != 1 else B
        B = B or 1
        rows = np.empty((B, cp.plan.in_elems), dtype=np.float64)
        for vars_, size, off in cp.plan.inputs:
            _, vals = items[vars_[0]]
            if vals.shape[0] not in (1, B):
                raise ValueError("soft evidence batch sizes disagree")
            rows[:, off:off + size] = vals
        return torch.from_numpy(rows).to(cp.device, dtype=cp.torch_dtype)

    def _run(self, cp: CompiledPlan, ev_states, soft=None):
        torch = require_cuda()
        if cp.n_ev == 0:
            B = int(ev_states.shape[0]) if ev_states is not None and hasattr(ev_states, "shape") else 1
            if soft is not None:
                B = int(soft.shape[0])
            out = torch.empty((max(B, 1), cp.out_elems), dtype=cp.torch_dtype, device=cp.device)
            return cp.run(None, out=out, soft=soft)
        if isinstance(ev_states, np.ndarray) or not hasattr(ev_states, "is_cuda"):
            ev = np.ascontiguousarray(np.asarray(ev_states, dtype=np.int32)).reshape(-1, cp.n_ev)
            for j, v in enumerate(cp.plan.ev_vars):
                if ev.shape[0] and (ev[:, j].min() < 0 or ev[:, j].max() >= self.cardinality[v]):
                    raise ValueError(f"evidence state index out of range for variable {v}")
            ev_t = torch.from_numpy(ev).to(cp.device)
        else:
            # evidence already on the GPU: it must live on the plan's device; its values are range-checked on the device
            # only when config.validate_device_evidence is on (a min/max reduction + one host sync per call) — the kernels
            # clamp out-of-range states into [0, card) for memory safety, so without the check a bad index yields a
            # plausible posterior for the clamped state instead of the ValueError the host-array path raises
            ev_t = ev_states
            if ev_t.device != cp.device:
                raise ValueError(f"evidence_states is on {ev_t.device}, the plan on {cp.device}")
            if config.validate_device_evidence and ev_t.numel():
                lo = ev_t.amin(dim=0).cpu().numpy()
                hi = ev_t.amax(dim=0).cpu().numpy()
                for j, v in enumerate(cp.plan.ev_vars):
                    if lo[j] < 0 or hi[j] >= self.cardinality[v]:
                        raise ValueError(f"evidence state index out of range for variable {v}")
        return cp.run(ev_t, soft=soft)

    @staticmethod
    def _warn_nan(values):
        if np.isnan(values).any():
            warnings.warn("invalid value encountered in divide", RuntimeWarning, stacklevel=3)


class VariableElimination(_Inference):
    """pgmpy.inference.VariableElimination (ExactInference.py:22-733) on the B200 engine."""

    def _virtual(self, virtual_evidence):
        """inference/base.py:256-299: add a binary child "__var" with CPD [p; 1-p] per soft-evidence item."""
        bn = self.model.copy()
        for cpd in virtual_evidence:
            if not isinstance(cpd, (TabularCPD, DiscreteFactor)) and not hasattr(cpd, "variables"):
                raise ValueError(
                    f"Virtual evidence should be an instance of TabularCPD or DiscreteFactor. Got: {type(cpd)}"
                )
            if len(cpd.variables) > 1:
                raise ValueError("Virtual evidence should be defined on individual variables.")
            var = cpd.variables[0]
            if var not in self.variables:
                raise ValueError("Evidence provided for variable which is not in the model")
            vals = np.asarray(cpd.values, dtype=np.float64).reshape(-1)
            if vals.size != self.cardinality[var]:
                raise ValueError(
                    "The number of states/cardinality for the evidence should"
                    " be same as the number of states/cardinality of the variable in the model"
                )
            new_var = "__" + str(var)
            bn.add_edge(var, new_var)
            bn.add_cpds(
                TabularCPD(
                    new_var, 2, np.vstack((vals, 1 - vals)), [var], [self.cardinality[var]],
                    state_names={new_var: [0, 1], var: list(cpd.state_names[var])},
                )
            )
        return bn

    def _plan(self, variables, ev_vars, joint, elimination_order, prune=True, reduce_max=False, soft_vars=()) -> CompiledPlan:
        order_key = tuple(elimination_order) if isinstance(elimination_order, (list, tuple)) else None
        key = ("ve", tuple(variables), tuple(ev_vars), joint, order_key, prune, reduce_max, tuple(soft_vars))
        cp = self._plans.get(key)
        if cp is None:
            if isinstance(self.model, JunctionTree):
                factors = [(tuple(f.variables), f.values, None) for f in self.model.get_factors()]
                plan = PL.compile_factor_ve_plan(
                    factors, self.cardinality, variables, ev_vars, joint=joint, normalize=True,
                    elimination_order=order_key, reduce_max=reduce_max, soft_vars=soft_vars,
                )
            else:
                plan = PL.compile_ve_plan(
                    self.model, variables, ev_vars, joint=joint, prune=prune, elimination_order=order_key,
                    reduce_max=reduce_max, soft_vars=soft_vars,
                )
            cp = self._compile(plan)
            self._plans[key] = cp
        return cp

    def query(
        self,
        variables,
        evidence=None,
        virtual_evidence=None,
        elimination_order="greedy",
        joint=True,
        show_progress=True,
    ):
        """Posterior over `variables` given `evidence` ({var: state name}). Returns a normalised
        DiscreteFactor (joint=True) or {var: DiscreteFactor} (joint=False), ExactInference.py:246-457.
        `elimination_order`: heuristic names are accepted for compatibility — the engine always uses its
        own min-fill order (order changes results only at ~1e-16); an explicit list is honoured."""
        evidence = self._check_query(variables, evidence)
        variables = list(variables)
        ev_vars = list(evidence)
        soft_vars, soft = (), None
        if virtual_evidence is not None and isinstance(self.model, DiscreteBayesianNetwork):
            # soft evidence is a per-evidence-set INPUT of the plan (one plan per signature, cached) — the factor the
            # reference gets by adding an observed binary child per query (inference/base.py:256-299)
            soft_vars = tuple(_soft_items(self, virtual_evidence))
        cp = self._plan(variables, ev_vars, joint, elimination_order, soft_vars=soft_vars)
        if soft_vars:
            soft = self._soft_rows(cp, virtual_evidence)
        states = self._states_of(ev_vars, [evidence])  # KeyError on unknown state names
        out = self._run(cp, states, soft).cpu().numpy()[0]
        self._warn_nan(out)
        if joint:
            shape = [self.cardinality[v] for v in variables]
            return self._to_factor(variables, out.reshape(shape))
        res = {}
        for seg in cp.plan.segments:
            v = seg.vars[0]
            res[v] = self._to_factor([v], out[seg.out_offset : seg.out_offset + seg.table.size])
        return res

    def query_batch(self, variables, evidence_vars, evidence_states, joint=True, elimination_order=None,
                    virtual_evidence=None):
        """One signature, B evidence sets. evidence_states: int32 [B, k] state INDICES (host array or CUDA
        tensor) in `evidence_vars` order; a host array is range-checked (ValueError), a CUDA tensor is clamped into
        range by the kernels unless `config.validate_device_evidence` is set. `virtual_evidence`: [(variable, likelihoods [B, card]), ...] — soft evidence
        that differs per evidence set (semantics of inference/base.py:256-299 per row).
        Returns a CUDA tensor [B, out_elems]: the joint over `variables` (row-major in the given order) or the
        concatenated per-variable marginals when joint=False."""
        self._check_query(variables, {v: None for v in evidence_vars})
        soft_vars = tuple(_soft_items(self, virtual_evidence)) if virtual_evidence else ()
        cp = self._plan(list(variables), list(evidence_vars), joint, elimination_order, soft_vars=soft_vars)
        B = int(np.shape(evidence_states)[0]) if len(evidence_vars) else None
        return self._run(cp, evidence_states, self._soft_rows(cp, virtual_evidence, B) if soft_vars else None)

    def marginals_plan(self, evidence_vars, variables=None) -> CompiledPlan:
        """One plan holding the VE-mode posterior of every unobserved variable (or of `variables`), each with its own
        pruned network exactly as `query([v], evidence)` would use — the reference's way of asking for all-variable
        marginals, as a single launch."""
        order = list(self.model.nodes())
        evset = set(evidence_vars)
        variables = [v for v in order if v not in evset] if variables is None else list(variables)
        self._check_query(variables, {v: None for v in evidence_vars})
        key = ("ve-multi", tuple(variables), tuple(evidence_vars))
        cp = self._plans.get(key)
        if cp is None:
            plan = PL.compile_ve_multi_plan(self.model, [[v] for v in variables], list(evidence_vars))
            cp = self._compile(plan)
            self._plans[key] = cp
        return cp

    def marginals_batch(self, evidence_vars, evidence_states, variables=None):
        """CUDA tensor [B, sum card]: per-variable VE-mode posteriors for B evidence sets (segments in
        `marginals_plan(...).plan.segments`)."""
        return self._run(self.marginals_plan(evidence_vars, variables), evidence_states)

    def query_batch_mixed(self, variables, evidence_rows, joint=True):
        """Mixed-evidence batch: `evidence_rows` is a list of {var: state name} dicts whose observed SETS may
        differ from row to row (what DiscreteBayesianNetwork.predict_probability feeds the reference one row at a
        time, pgmpy/models/DiscreteBayesianNetwork.py:973-989). Rows are bucketed by evidence-variable signature,
        every bucket runs as one batched plan, results come back in row order as a CUDA tensor [B, out_elems]."""
        torch = require_cuda()
        buckets: Dict[tuple, List[int]] = {}
        for i, row in enumerate(evidence_rows):
            buckets.setdefault(tuple(sorted(row, key=str)), []).append(i)
        jobs = []
        for sig, idx in buckets.items():
            ev_vars = list(sig)
            self._check_query(variables, {v: None for v in ev_vars})
            cp = self._plan(list(variables), ev_vars, joint, None)
            jobs.append((cp, self._states_of(ev_vars, [evidence_rows[i] for i in idx]), idx))
        if not jobs:
            return None
        cp0 = jobs[0][0]
        # one upload of all evidence, ONE call across the C-ABI for all buckets (pgx_run_batch_multi), one scatter back
        # into row order: per bucket the host path (~15 us of Python + ctypes, a copy, an indexing kernel) was longer
        # than the bucket's kernel
        from .engine import MultiRun

        flat, spans, order = [], [], []
        for cp, states, idx in jobs:
            ev = np.ascontiguousarray(np.asarray(states, dtype=np.int32)).reshape(len(idx), cp.n_ev)
            for j, v in enumerate(cp.plan.ev_vars):
                if ev.shape[0] and (ev[:, j].min() < 0 or ev[:, j].max() >= self.cardinality[v]):
                    raise ValueError(f"evidence state index out of range for variable {v}")
            spans.append((sum(a.size for a in flat), ev.shape))
            flat.append(ev.ravel())
            order.extend(idx)
        with torch.cuda.device(cp0.device):
            ev_all = torch.from_numpy(np.concatenate(flat) if flat else np.zeros(0, np.int32)).to(cp0.device)
            res_all = torch.empty((len(order), cp0.out_elems), dtype=cp0.torch_dtype, device=cp0.device)
            mr_jobs, lo = [], 0
            for (cp, _, idx), (off, shape) in zip(jobs, spans):
                ev_t = ev_all[off:off + shape[0] * shape[1]].view(shape) if cp.n_ev else None
                mr_jobs.append((cp, ev_t, res_all[lo:lo + len(idx)]))
                lo += len(idx)
            MultiRun(mr_jobs).run()
            if order == list(range(len(order))):
                return res_all
            out = torch.empty_like(res_all)
            out[torch.as_tensor(order, device=cp0.device)] = res_all
        return out

    # ---- max-product queries ---------------------------------------------
    def _argmax_rows(self, table):
        """Row-wise argmax on the device (pgx_argmax_rows)."""
        import ctypes as C

        from . import _native as N

        torch = require_cuda()
        out = torch.empty((table.shape[0],), dtype=torch.int32, device=table.device)
        N.check(N.load().pgx_argmax_rows(
            N.PGX_F64 if table.dtype == torch.float64 else N.PGX_F32, C.c_void_p(table.data_ptr()), table.shape[1],
            table.shape[0], C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(table.device).cuda_stream)))
        return out

    def _decode(self, variables, flat_index) -> dict:
        """DiscreteFactor.assignment for one flat index of the joint over `variables` (row-major)."""
        res = {}
        for v in reversed(list(variables)):
            c = self.cardinality[v]
            res[v] = self.states[v][int(flat_index % c)]
            flat_index //= c
        return {v: res[v] for v in variables}

    def max_marginal(self, variables=None, evidence=None, elimination_order="MinFill", show_progress=True):
        """ExactInference.py:459-526: eliminate every other variable with MAX, normalise the table over `variables`
        (the reference's _variable_elimination normalises joint results of Bayesian networks, :226-229) and return
        its largest entry. With no `variables` the reference multiplies all factors and takes the maximum of the
        full joint (:185-190) — the probability of the most probable explanation; we max-eliminate everything."""
        variables = list(variables) if variables else []
        evidence = self._check_query(variables, evidence, allow_empty=True)
        torch = require_cuda()
        if not variables:
            # evidence is ignored on this path by the reference (early return before the factors are reduced)
            nodes = list(self.model.nodes()) if not isinstance(self.model, JunctionTree) else sorted(self.variables, key=str)
            return float(self._run_unnormalized_max(nodes))
        ev_vars = list(evidence)
        cp = self._plan(variables, ev_vars, True, None, reduce_max=True)
        out = self._run(cp, self._states_of(ev_vars, [evidence]))
        return float(out.max().item())

    def _run_unnormalized_max(self, nodes):
        """max over the full joint of prod(all factors) — one max-elimination plan whose last table is emitted raw."""
        key = ("mpe-value",)
        cp = self._plans.get(key)
        if cp is None:
            if isinstance(self.model, JunctionTree):
                factors = [(tuple(f.variables), f.values, None) for f in self.model.get_factors()]
            else:
                factors = [(tuple(c.variables), c.values, None) for c in self.model.get_cpds()]
            plan = PL.compile_factor_ve_plan(factors, self.cardinality, [nodes[0]], [], joint=True, normalize=False,
                                             reduce_max=True)
            cp = self._compile(plan)
            self._plans[key] = cp
        out = self._run(cp, np.zeros((1, 0), dtype=np.int32))
        return out.max().item()

    def map_query(self, variables=None, evidence=None, virtual_evidence=None, elimination_order="MinFill",
                  show_progress=True):
        """ExactInference.py:528-624: sum-product joint over `variables` given the evidence, then the argmax
        assignment (state names). `variables=None` means every unobserved variable."""
        evidence = dict(evidence) if evidence is not None else {}
        if virtual_evidence is not None and isinstance(self.model, DiscreteBayesianNetwork):
            sub = type(self)(self._virtual(virtual_evidence), dtype=self.dtype)
            virt = {"__" + str(c.variables[0]): 0 for c in virtual_evidence}
            return sub.map_query(variables, {**evidence, **virt}, None, elimination_order, show_progress)
        if not variables:
            if isinstance(self.model, DiscreteBayesianNetwork):
                # every unobserved variable (nothing is pruned then): junction-tree max-product with back-pointers
                # instead of the full joint table (BeliefPropagation.mpe_batch)
                bp = getattr(self, "_mpe_bp", None)
                if bp is None:
                    bp = self._mpe_bp = BeliefPropagation(self.model, dtype=self.dtype)
                return bp.map_query(None, evidence)
            order = sorted(self.variables, key=str)
            variables = [v for v in order if v not in evidence]
        variables = list(variables)
        evidence = self._check_query(variables, evidence)
        ev_vars = list(evidence)
        joint = self._joint_for_map(variables, ev_vars, self._states_of(ev_vars, [evidence]))
        idx = int(self._argmax_rows(joint)[0].item())
        return self._decode(variables, idx)

    def _joint_for_map(self, variables, ev_vars, states):
        cp = self._plan(variables, ev_vars, True, None)
        return self._run(cp, states)

    def map_query_batch(self, variables, evidence_vars, evidence_states):
        """Batched map_query: int32 CUDA tensor [B, len(variables)] of state indices of the MAP assignment."""
        torch = require_cuda()
        variables = list(variables)
        self._check_query(variables, {v: None for v in evidence_vars})
        joint_size = 1
        for v in variables:
            joint_size *= self.cardinality[v]
        if (isinstance(self.model, DiscreteBayesianNetwork) and joint_size > (1 << 22)
                and set(variables) | set(evidence_vars) == set(self.variables)):
            # every unobserved variable is asked for (what predict() does for rows with many missing columns) and the
            # joint table would not fit: max-product with back-pointers gives the same argmax (ties aside)
            bp = getattr(self, "_mpe_bp", None)
            if bp is None:
                bp = self._mpe_bp = BeliefPropagation(self.model, dtype=self.dtype)
            cols, assign = bp.mpe_batch(list(evidence_vars), evidence_states)
            order = torch.as_tensor([cols.index(v) for v in variables], device=assign.device)
            return assign.index_select(1, order).contiguous()
        joint = self._joint_for_map(variables, list(evidence_vars), evidence_states)
        flat = self._argmax_rows(joint).to(torch.int64)
        cols = []
        for v in reversed(variables):
            c = self.cardinality[v]
            cols.append((flat % c).to(torch.int32))
            flat = flat // c
        return torch.stack(cols[::-1], dim=1)

    def induced_graph(self, elimination_order):
        """Induced graph of running variable elimination in the given order (host graph code,
        ExactInference.py:626-691): a networkx.Graph whose edges join every pair of variables that share a factor
        at some point of the elimination."""
        import itertools

        import networkx as nx

        if isinstance(self.model, JunctionTree):
            scopes = [list(f.variables) for f in self.model.get_factors()]
        else:
            scopes = [list(c.variables) for c in self.model.get_cpds()]
        if set(elimination_order) != set(self.variables):
            raise ValueError("Set of variables in elimination order different from variables in model")
        working = {v: [sc for sc in scopes if v in sc] for v in self.variables}
        cliques = {tuple(sc) for sc in scopes}
        eliminated = set()
        for var in elimination_order:
            factors = [f for f in working[var] if not set(f) & eliminated]
            phi = set(itertools.chain(*factors)) - {var}
            cliques.add(tuple(phi))
            del working[var]
            for v in phi:
                working[v].append(list(phi))
            eliminated.add(var)
        edges = itertools.chain(*[itertools.combinations(c, 2) for c in cliques if len(c) > 1])
        return nx.Graph(edges)

    def induced_width(self, elimination_order):
        """Size of the largest clique of the induced graph minus one (ExactInference.py:693-733)."""
        import networkx as nx

        return max(len(c) for c in nx.find_cliques(self.induced_graph(elimination_order))) - 1


class BeliefPropagation(_Inference):
    """pgmpy.inference.BeliefPropagation (ExactInference.py:736-1317) on the B200 engine.

    The junction tree is our min-fill tree (the reference's own builder blows up on alarm already,
    SURVEY.md §0 fact 5) unless a JunctionTree is passed in, which is then used as given
    (ExactInference.py:742-745)."""

    def __init__(self, model, dtype: Optional[str] = None):
        super().__init__(model, dtype)
        if self._unsupported:
            return
        if isinstance(self.model, JunctionTree):
            self.junction_tree = self.model
            self._jt = PL.JTStructure.from_junction_tree(self.model)
        else:
            self._jt = PL.JTStructure.from_model(self.model)
            self.junction_tree = None
        self.clique_beliefs = {}
        self.sepset_beliefs = {}

    def get_cliques(self):
        return list(self._jt.cliques)

    def get_clique_beliefs(self):
        return self.clique_beliefs

    def get_sepset_beliefs(self):
        return self.sepset_beliefs

    def _jt_plan(self, ev_vars, variables=None, emit_beliefs=False, soft_vars=(), specialize=False) -> CompiledPlan:
        key = ("jt", tuple(ev_vars), None if variables is None else tuple(variables), emit_beliefs, tuple(soft_vars), bool(specialize))
        cp = self._plans.get(key)
        if cp is None:
            plan = PL.compile_jt_plan(self._jt, ev_vars, variables, emit_beliefs=emit_beliefs, soft_vars=soft_vars,
                                      objective="flops" if specialize else "bytes")
            cp = self._compile(plan)
            if specialize:
                try:
                    cp.specialize()
                except Exception:  # max / input plans, no libnvrtc: the other kernels serve the plan
                    pass
            self._plans[key] = cp
        return cp

    def calibrate(self):
        """Calibrated (un-normalised) clique and sepset beliefs without evidence, ExactInference.py:897-945.
        One collect + one distribute pass on the GPU reaches the fixed point the reference iterates to."""
        self._calibrate(reduce_max=False)

    def _calibrate(self, reduce_max: bool):
        key = ("jt-beliefs", reduce_max)
        cp = self._plans.get(key)
        if cp is None:
            cp = self._compile(PL.compile_jt_plan(self._jt, [], None, emit_beliefs=True, reduce_max=reduce_max))
            self._plans[key] = cp
        out = self._run(cp, np.zeros((1, 0), dtype=np.int32)).cpu().numpy()[0]
        self.clique_beliefs = {}
        self.sepset_beliefs = {}
        n = len(self._jt.cliques)
        for i, seg in enumerate(cp.plan.segments):
            vals = out[seg.out_offset : seg.out_offset + seg.table.size]
            f = self._to_factor(list(seg.vars), vals.reshape([self.cardinality[v] for v in seg.vars]))
            if i < n:
                self.clique_beliefs[self._jt.cliques[i]] = f
        k = n
        for i in range(n):
            p = self._jt.parent[i]
            if p >= 0:
                seg = cp.plan.segments[k]
                k += 1
                vals = out[seg.out_offset : seg.out_offset + seg.table.size]
                f = self._to_factor(list(seg.vars), vals.reshape([self.cardinality[v] for v in seg.vars]))
                self.sepset_beliefs[frozenset((self._jt.cliques[i], self._jt.cliques[p]))] = f

    def query(self, variables, evidence=None, virtual_evidence=None, joint=True, show_progress=True):
        """ExactInference.py:1117-1220. BP mode = all factors, no pruning: single
        variables come from the junction-tree plan, joints over several variables from an un-pruned
        elimination plan (the reference's out-of-clique query, :1047-1111, computes the same function)."""
        evidence = self._check_query(variables, evidence)
        variables = list(variables)
        ev_vars = list(evidence)
        states = self._states_of(ev_vars, [evidence])
        soft_vars = ()
        if virtual_evidence is not None and isinstance(self.model, DiscreteBayesianNetwork):
            soft_vars = tuple(_soft_items(self, virtual_evidence))  # likelihood vectors = input tables of the plan
        if len(variables) == 1 or not joint:
            cp = self._jt_plan(ev_vars, variables, soft_vars=soft_vars)
            out = self._run(cp, states, self._soft_rows(cp, virtual_evidence) if soft_vars else None).cpu().numpy()[0]
            self._warn_nan(out)
            res = {}
            for seg in cp.plan.segments:
                v = seg.vars[0]
                res[v] = self._to_factor([v], out[seg.out_offset : seg.out_offset + seg.table.size])
            return res[variables[0]] if joint else res
        key = ("bp-joint", tuple(variables), tuple(ev_vars), soft_vars)
        cp = self._plans.get(key)
        if cp is None:
            factors = [(c, p, None) for c, p in zip(self._jt.cliques, self._jt.potentials)]
            plan = PL.compile_factor_ve_plan(factors, self.cardinality, variables, ev_vars, joint=True, normalize=True,
                                             soft_vars=soft_vars)
            cp = self._compile(plan)
            self._plans[key] = cp
        out = self._run(cp, states, self._soft_rows(cp, virtual_evidence) if soft_vars else None).cpu().numpy()[0]
        self._warn_nan(out)
        return self._to_factor(variables, out.reshape([self.cardinality[v] for v in variables]))

    def max_calibrate(self):
        """ExactInference.py:947-995: max-calibrated clique and sepset beliefs (messages reduced with max)."""
        self._calibrate(reduce_max=True)

    def map_query(self, variables=None, evidence=None, virtual_evidence=None, show_progress=True):
        """ExactInference.py:1222-1317: argmax assignment of the (un-pruned, BP-mode) joint over `variables`."""
        evidence = dict(evidence) if evidence is not None else {}
        if virtual_evidence is not None and isinstance(self.model, DiscreteBayesianNetwork):
            ve = VariableElimination(self.model, dtype=self.dtype)
            sub = BeliefPropagation(ve._virtual(virtual_evidence), dtype=self.dtype)
            virt = {"__" + str(c.variables[0]): 0 for c in virtual_evidence}
            return sub.map_query(variables, {**evidence, **virt}, None, show_progress)
        if not variables:
            # every unobserved variable: max-product with back-pointers (the joint table the reference maximises would
            # have prod(card) entries)
            evidence = self._check_query([], evidence, allow_empty=True)
            ev_vars = list(evidence)
            columns, assign = self.mpe_batch(ev_vars, self._states_of(ev_vars, [evidence]))
            row = assign[0].cpu().numpy()
            return {v: self.states[v][int(s)] for v, s in zip(columns, row)}
        joint = self.query(list(variables), evidence=evidence, joint=True)
        torch = require_cuda()
        t = torch.from_numpy(np.ascontiguousarray(joint.values.reshape(1, -1))).cuda()
        idx = int(VariableElimination._argmax_rows(self, t)[0].item())
        return VariableElimination._decode(self, list(variables), idx)

    def mpe_plan(self, evidence_vars, soft_vars=()):
        """(compiled max-product plan with its traceback attached, the variable of each output column)."""
        key = ("jt-mpe", tuple(evidence_vars), tuple(soft_vars))
        hit = self._plans.get(key)
        if hit is None:
            plan, trace, columns = PL.compile_jt_mpe_plan(self._jt, list(evidence_vars), soft_vars=soft_vars)
            cp = self._compile(plan)
            cp.set_trace(trace)
            hit = (cp, columns)
            self._plans[key] = hit
        return hit

    def mpe_batch(self, evidence_vars, evidence_states, virtual_evidence=None):
        """Most probable explanation of ALL unobserved variables for B evidence sets: max-product
        on the junction tree with back-pointers instead of the reference's argmax over the full joint table
        (ExactInference.py:609-612, :1222-1317 — 10^15 entries on alarm). Returns (variables, int32 CUDA tensor
        [B, len(variables)] of state indices). Among exactly tied maxima the first in clique order wins, which need not
        be the reference's first-in-joint-order choice; the joint probability of the answer is the same."""
        torch = require_cuda()
        self._check_query([], {v: None for v in evidence_vars}, allow_empty=True)
        soft_vars = tuple(_soft_items(self, virtual_evidence)) if virtual_evidence else ()
        cp, columns = self.mpe_plan(evidence_vars, soft_vars)
        ev_t = None
        if cp.n_ev:
            if isinstance(evidence_states, np.ndarray) or not hasattr(evidence_states, "is_cuda"):
                ev = np.ascontiguousarray(np.asarray(evidence_states, dtype=np.int32)).reshape(-1, cp.n_ev)
                for j, v in enumerate(cp.plan.ev_vars):
                    if ev.shape[0] and (ev[:, j].min() < 0 or ev[:, j].max() >= self.cardinality[v]):
                        raise ValueError(f"evidence state index out of range for variable {v}")
                ev_t = torch.from_numpy(ev).to(cp.device)
            else:
                ev_t = evidence_states
        soft = self._soft_rows(cp, virtual_evidence, int(ev_t.shape[0]) if ev_t is not None else None) if soft_vars else None
        return columns, cp.run_mpe(ev_t, soft)

    def marginals_plan(self, evidence_vars, variables=None, soft_vars=(), specialize=False) -> CompiledPlan:
        """Compiled all-marginals plan for one evidence-variable signature (bench / batched callers).
        `specialize=True`: the caller expects ~10^8 or more evidence sets on this signature — the plan variant with the
        fewest multiply-adds is chosen (planner.plan_flops) and compiled into its own straight-line kernel right away
        (CompiledPlan.specialize, about a second)."""
        self._check_query(variables or [], {v: None for v in evidence_vars}, allow_empty=True)
        return self._jt_plan(list(evidence_vars), variables, soft_vars=tuple(soft_vars), specialize=specialize)

    def marginals_batch(self, evidence_vars, evidence_states, variables=None, virtual_evidence=None):
        """Posterior marginals of every unobserved variable (or `variables`) for B evidence sets:
        CUDA tensor [B, sum card]; column layout in `marginals_plan(...).plan.segments`.
        `virtual_evidence`: [(variable, likelihoods [B, card]), ...] — per-evidence-set soft evidence."""
        soft_vars = tuple(_soft_items(self, virtual_evidence)) if virtual_evidence else ()
        cp = self.marginals_plan(evidence_vars, variables, soft_vars)
        B = int(np.shape(evidence_states)[0]) if len(evidence_vars) else None
        return self._run(cp, evidence_states, self._soft_rows(cp, virtual_evidence, B) if soft_vars else None)
