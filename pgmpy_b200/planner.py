"""Plan compiler: model (+ query/evidence signature) -> static contraction plan (pgmpy_b200.plan.Plan).

Runs once per model / per signature on the host. Two plan families, matching the two closed forms the
reference computes:

  * VE mode  (`compile_ve_plan`, prune=True): normalise( sum_{K \\ (Q u E)} prod_{v in K} CPT'_v[E=e] ), K = kept
    nodes of Inference._prune_bayesian_model (pgmpy/inference/base.py:154-212), CPT'_v = CPT with the
    dropped parents summed out and columns renormalised (pgmpy/factors/discrete/CPD.py:483-524).
    Steps follow the elimination loop of VariableElimination._variable_elimination
    (pgmpy/inference/ExactInference.py:200-229) with OUR min-fill order, each step fused product+sum-out.
  * BP mode  (`compile_jt_plan`, or compile_ve_plan(prune=False)): all factors, no pruning — what
    BeliefPropagation.calibrate/query computes (ExactInference.py:854-895, :997-1111). The junction tree
    is ours (min-fill); message passing is a two-pass collect/distribute schedule whose messages are
    fused product+sum-out steps over the clique potential and the incoming messages.
"""
from __future__ import annotations

from typing import Dict, Hashable, List, Optional, Sequence, Tuple

import numpy as np

from . import graph as G
from .factors import DiscreteFactor, TabularCPD
from .models import DiscreteBayesianNetwork, JunctionTree
from .plan import Plan, PlanBuilder, Table


# ---------------------------------------------------------------------------------------------
# host-side constant folding (batch-invariant, once per plan)
# ---------------------------------------------------------------------------------------------
def _product_tables(tables: Sequence[Tuple[Sequence[Hashable], np.ndarray]], out_vars: Sequence[Hashable]) -> np.ndarray:
    """Dense product of constant tables over `out_vars` (used to build clique potentials once per model,
    as the reference does at pgmpy/models/DiscreteMarkovNetwork.py:598-629)."""
    idx = {v: i for i, v in enumerate(out_vars)}
    dims = {}
    args = []
    for vars_, vals in tables:
        args += [vals, [idx[v] for v in vars_]]
        for v, d in zip(vars_, np.shape(vals)):
            dims[v] = d
    if not args:
        return np.ones(())
    present = [v for v in out_vars if v in dims]
    prod = np.einsum(*args, [idx[v] for v in present], optimize=False)
    # broadcast over clique variables that no assigned table mentions
    shape = [dims.get(v, 1) for v in out_vars]
    return prod.reshape(shape)


def marginalized_cpd(cpd: TabularCPD, drop: Sequence[Hashable]) -> Tuple[List[Hashable], np.ndarray]:
    """CPT with parents `drop` summed out and columns renormalised — what the reference's pruning does
    to CPDs whose scope leaves the kept set (base.py:203-208 -> CPD.py:483-524: sum, then
    `cpd / cpd.sum(axis=0)` on the 2-D view)."""
    keep_axes = [i for i, v in enumerate(cpd.variables) if v not in drop]
    vals = np.einsum(cpd.values, list(range(len(cpd.variables))), keep_axes)
    two_d = vals.reshape(vals.shape[0], -1)
    two_d = two_d / two_d.sum(axis=0)
    return [cpd.variables[i] for i in keep_axes], two_d.reshape(vals.shape)


# ---------------------------------------------------------------------------------------------
# junction tree of a Bayesian network (ours)
# ---------------------------------------------------------------------------------------------
def jt_structure(model: DiscreteBayesianNetwork):
    nodes = model.nodes()
    rank = {v: i for i, v in enumerate(nodes)}
    card = model.get_cardinality()
    parents = {n: model.get_parents(n) for n in nodes}
    adj = G.moral_graph(parents)
    cliques, edges = G.junction_tree(adj, card, rank)
    return cliques, edges


def assign_factors(cliques: Sequence[Tuple[Hashable, ...]], scopes: Sequence[Sequence[Hashable]], card) -> List[int]:
    """Each factor goes to the smallest clique containing its scope (every factor used exactly once)."""
    sets = [set(c) for c in cliques]
    sizes = [int(np.prod([card[v] for v in c], dtype=np.int64)) for c in cliques]
    by_var: Dict[Hashable, List[int]] = {}
    for i, c in enumerate(cliques):
        for v in c:
            by_var.setdefault(v, []).append(i)
    out = []
    for sc in scopes:
        sc = list(sc)
        cand = by_var[sc[0]]
        best = None
        for i in cand:
            if all(v in sets[i] for v in sc):
                if best is None or sizes[i] < sizes[best]:
                    best = i
        if best is None:
            raise ValueError(f"no clique contains factor scope {sc}")
        out.append(best)
    return out


def build_junction_tree(model: DiscreteBayesianNetwork) -> JunctionTree:
    """Min-fill junction tree with clique potentials = product of the assigned CPDs (ones where none),
    the same object shape BeliefPropagation.__init__ accepts (ExactInference.py:742-745)."""
    cliques, edges = jt_structure(model)
    card = model.get_cardinality()
    states = model.states
    cpds = model.get_cpds()
    owner = assign_factors(cliques, [c.variables for c in cpds], card)
    jt = JunctionTree()
    jt.add_nodes_from(cliques)
    for a, b in edges:
        if set(cliques[a]) & set(cliques[b]):
            jt.add_edge(cliques[a], cliques[b])
    for i, c in enumerate(cliques):
        mine = [(cpd.variables, cpd.values) for cpd, o in zip(cpds, owner) if o == i]
        dims = [card[v] for v in c]
        vals = np.ones(dims) * _product_tables(mine, c) if mine else np.ones(dims)
        jt.factors.append(DiscreteFactor(list(c), dims, vals, {v: states[v] for v in c}))
    return jt


# ---------------------------------------------------------------------------------------------
# VE plans
# ---------------------------------------------------------------------------------------------
def _pruned_factors(model: DiscreteBayesianNetwork, variables, evidence_vars, prune: bool, soft_vars=()):
    """`soft_vars`: variables carrying soft (virtual) evidence. The reference gives each an observed binary child
    (pgmpy/inference/base.py:256-299) BEFORE it prunes, so for the kept set they count as parents of evidence."""
    nodes = model.nodes()
    if prune:
        parents = {n: list(model.get_parents(n)) for n in nodes}
        children = {n: list(model.get_children(n)) for n in nodes}
        virt = []
        for v in soft_vars:
            node = ("__soft__", v)
            parents[node] = [v]
            children[node] = []
            children[v] = children[v] + [node]
            virt.append(node)
        kept = G.prune_nodes(parents, children, list(variables), list(evidence_vars) + virt)
        kept = set(kept) - set(virt)
    else:
        kept = set(nodes)
    factors = []
    for v in nodes:
        if v not in kept:
            continue
        cpd = model.get_cpds(v)
        drop = [p for p in cpd.variables[1:] if p not in kept]
        if drop:
            vars_, vals = marginalized_cpd(cpd, drop)
        else:
            vars_, vals = list(cpd.variables), cpd.values
        factors.append((tuple(vars_), vals, ("cpd", v, tuple(drop))))
    return kept, factors


def compile_factor_ve_plan(
    factors: Sequence[Tuple[Sequence[Hashable], np.ndarray, object]],
    card: Dict[Hashable, int],
    variables: Sequence[Hashable],
    evidence_vars: Sequence[Hashable],
    joint: bool = True,
    normalize: bool = True,
    elimination_order: Optional[Sequence[Hashable]] = None,
    rank: Optional[Dict[Hashable, int]] = None,
    meta: Optional[dict] = None,
    reduce_max: bool = False,
    builder: Optional[PlanBuilder] = None,
    soft_vars: Sequence[Hashable] = (),
) -> Plan:
    """Sum-product variable elimination over an explicit factor list.

    `factors` = (scope, values, key). Factors whose whole scope is observed are dropped, as both
    reference paths do (greedy: ExactInference.py:383-384; classic: scalar factors fall out of
    `working_factors`, :53-66). Output: the joint over `variables` in the caller's order
    (ExactInference.py:404-414), or one marginal per variable when joint=False (:424-433)."""
    variables = list(variables)
    ev = list(evidence_vars)
    evset = set(ev)
    b = builder if builder is not None else PlanBuilder(card, ev)
    work: List[Table] = []
    for scope, vals, key in factors:
        if all(v in evset for v in scope):
            continue
        work.append(b.add_const(scope, vals, key=key if builder is not None else None))
    for v in soft_vars:
        # soft evidence = one more factor over v whose values arrive per evidence set (the reduced CPD of the observed
        # binary child the reference adds, inference/base.py:256-299)
        if v in evset:
            raise ValueError(f"soft evidence on the observed variable {v}")
        work.append(b.add_input([v]))
    free_scopes = [[v for v in t.vars if v not in evset] for t in work]
    present = set(v for sc in free_scopes for v in sc)
    for q in variables:
        if q not in present:
            raise ValueError(f"query variable {q} not in any factor")
    if elimination_order is None:
        adj = G.interaction_graph(free_scopes)
        if rank is None:
            rank = {v: i for i, v in enumerate(sorted(adj, key=str))}
        order, _ = G.min_fill_order(adj, card, keep=variables, rank={v: rank.get(v, 0) for v in adj})
    else:
        order = [v for v in elimination_order if v in present and v not in variables]
        missing = present - set(order) - set(variables)
        if missing:
            raise ValueError(f"elimination order misses variables: {sorted(map(str, missing))}")
    level = 0
    for var in order:
        touching = [t for t in work if var in t.vars]
        if not touching:
            continue
        work = [t for t in work if var not in t.vars]
        scope: List[Hashable] = []
        biggest = max(touching, key=lambda t: t.size)
        for t in [biggest] + touching:
            for v in t.vars:
                if v not in evset and v != var and v not in scope:
                    scope.append(v)
        out = b.contract(touching, scope, level=level, reduce_max=reduce_max)
        level += 1
        work.append(out)
    # every remaining factor has scope within `variables`
    joint_t = b.contract(work, variables, level=level, reduce_max=reduce_max)
    if joint:
        b.emit(joint_t, normalize, variables)
    else:
        for q in variables:
            if len(variables) == 1:
                b.emit(joint_t, normalize, [q])
            else:
                m = b.contract([joint_t], [q], level=level + 1)
                b.emit(m, normalize, [q])
    m = {"mode": "ve-max" if reduce_max else "ve", "variables": tuple(variables), "evidence_vars": tuple(ev), "joint": joint,
         "order": tuple(order), "soft_vars": tuple(soft_vars)}
    m.update(meta or {})
    if builder is not None:
        return m  # the caller finalizes the shared builder
    return b.finalize(m)


def compile_ve_plan(
    model: DiscreteBayesianNetwork,
    variables: Sequence[Hashable],
    evidence_vars: Sequence[Hashable] = (),
    joint: bool = True,
    prune: bool = True,
    elimination_order: Optional[Sequence[Hashable]] = None,
    reduce_max: bool = False,
    soft_vars: Sequence[Hashable] = (),
) -> Plan:
    """reduce_max=True eliminates with max instead of sum (DiscreteFactor.maximize): the table the reference's
    max_marginal takes its maximum of (ExactInference.py:459-526). `soft_vars`: variables with soft (virtual)
    evidence; their likelihood vectors are input tables of the plan (Plan.inputs), one row per evidence set."""
    kept, factors = _pruned_factors(model, variables, evidence_vars, prune, soft_vars)
    ev = [v for v in evidence_vars]  # all evidence variables survive pruning (base.py:192-194)
    rank = {v: i for i, v in enumerate(model.nodes())}
    return compile_factor_ve_plan(
        factors,
        model.get_cardinality(),
        variables,
        ev,
        joint=joint,
        normalize=True,
        elimination_order=elimination_order,
        rank=rank,
        meta={"kept": tuple(sorted(kept, key=lambda v: rank[v])), "prune": prune},
        reduce_max=reduce_max,
        soft_vars=soft_vars,
    )


def compile_ve_multi_plan(
    model: DiscreteBayesianNetwork, queries: Sequence[Sequence[Hashable]], evidence_vars: Sequence[Hashable] = ()
) -> Plan:
    """Several VE-mode queries that share one evidence-variable set, as ONE plan (one launch): every query keeps its
    own pruned factor set (pruning depends on the query variables), identical (renormalised) CPTs
    are stored once, the per-query elimination chains are independent and run level-parallel; one normalised output
    segment per query. This is how "all-variable marginals" is asked of the reference (one query([v], e) per
    variable) without paying one launch per variable."""
    ev = list(evidence_vars)
    card = model.get_cardinality()
    rank = {v: i for i, v in enumerate(model.nodes())}
    b = PlanBuilder(card, ev)
    metas = []
    for variables in queries:
        _, factors = _pruned_factors(model, list(variables), ev, True)
        metas.append(compile_factor_ve_plan(factors, card, list(variables), ev, joint=True, normalize=True, rank=rank, builder=b))
    return b.finalize({"mode": "ve-multi", "evidence_vars": tuple(ev), "queries": tuple(tuple(q) for q in queries),
                       "orders": tuple(m["order"] for m in metas)})


# ---------------------------------------------------------------------------------------------
# junction-tree (BP mode) plans
# ---------------------------------------------------------------------------------------------
class JTStructure:
    """Rooted junction tree + clique potentials, built once per model and shared by all signatures.

    A clique potential is kept in two forms: `factors[i]`, the list of (scope, table) assigned to clique i — for a
    Bayesian network the CPDs themselves, never multiplied out — and `potentials[i]`, their dense product over the
    clique (what the reference builds at pgmpy/models/DiscreteMarkovNetwork.py:598-629), computed on first use. The
    big cliques of diabetes / munin (1e5 .. 2.7e6 entries) hold one 3-variable CPD or none at all: plans that multiply
    the factors into the messages (`compile_jt_plan(factorized=True)`) never touch a clique-sized constant table."""

    def __init__(self, cliques, edges, potentials, card, states, factors=None):
        self.cliques = [tuple(c) for c in cliques]
        self._potentials = potentials  # list of ndarray shaped by clique dims, or None until first use
        self.factors = factors if factors is not None else [[(tuple(c), p)] for c, p in zip(self.cliques, potentials)]
        self.card = dict(card)
        self.states = states
        n = len(self.cliques)
        self.nb: List[List[int]] = [[] for _ in range(n)]
        for a, b in edges:
            self.nb[a].append(b)
            self.nb[b].append(a)
        self.edges = [tuple(e) for e in edges]
        self.root = self._centre()
        self.parent = [-1] * n
        self.depth = [0] * n
        self.pre: List[int] = []
        stack = [self.root]
        seen = {self.root}
        while stack:
            x = stack.pop()
            self.pre.append(x)
            for y in self.nb[x]:
                if y not in seen:
                    seen.add(y)
                    self.parent[y] = x
                    self.depth[y] = self.depth[x] + 1
                    stack.append(y)
        if len(self.pre) != n:
            raise ValueError("junction tree is not connected")
        self.children = [[y for y in self.nb[x] if self.parent[y] == x] for x in range(n)]
        self.height = [0] * n
        for x in reversed(self.pre):
            for y in self.children[x]:
                self.height[x] = max(self.height[x], self.height[y] + 1)

    @property
    def potentials(self):
        if self._potentials is None:
            pots = []
            for c, mine in zip(self.cliques, self.factors):
                dims = [self.card[v] for v in c]
                pots.append(np.ones(dims) * _product_tables(mine, c) if mine else np.ones(dims))
            self._potentials = pots
        return self._potentials

    def _centre(self) -> int:
        n = len(self.cliques)
        if n == 1:
            return 0

        def far(src):
            dist = {src: 0}
            order = [src]
            for x in order:
                for y in self.nb[x]:
                    if y not in dist:
                        dist[y] = dist[x] + 1
                        order.append(y)
            last = order[-1]
            return last, dist

        a, _ = far(0)
        b, da = far(a)
        _, db = far(b)
        diam = da[b]
        best = min(range(n), key=lambda x: (max(da.get(x, n), db.get(x, n)), x))
        return best

    @classmethod
    def from_model(cls, model: DiscreteBayesianNetwork) -> "JTStructure":
        cliques, edges = jt_structure(model)
        card = model.get_cardinality()
        cpds = model.get_cpds()
        owner = assign_factors(cliques, [c.variables for c in cpds], card)
        factors = [[] for _ in cliques]
        for cpd, o in zip(cpds, owner):
            factors[o].append((tuple(cpd.variables), np.asarray(cpd.values, dtype=np.float64)))
        return cls(cliques, edges, None, card, model.states, factors=factors)

    @classmethod
    def from_junction_tree(cls, jt: JunctionTree) -> "JTStructure":
        cliques = jt.nodes()
        index = {c: i for i, c in enumerate(cliques)}
        edges = [(index[tuple(u)], index[tuple(v)]) for u, v in jt.edges()]
        card = jt.get_cardinality()
        pots = []
        for c in cliques:
            f = jt.get_factors(c)
            perm = [f.variables.index(v) for v in c]
            pots.append(np.ascontiguousarray(np.transpose(f.values, perm)))
        # join a forest with empty-sepset edges
        comp = list(range(len(cliques)))

        def cf(i):
            while comp[i] != i:
                comp[i] = comp[comp[i]]
                i = comp[i]
            return i

        for a, b in edges:
            comp[cf(a)] = cf(b)
        reps = sorted({cf(i) for i in range(len(cliques))})
        for k in range(1, len(reps)):
            edges.append((reps[0], reps[k]))
        return cls(cliques, edges, pots, card, jt.states)


def plan_cost(plan: Plan) -> float:
    """Estimated ns per evidence set (PlanBuilder's cost model: HBM time of the work-table traffic + issue time of the
    operand loads). Used to choose between plan variants of the same query."""
    loads = plan.operand_loads()
    nbytes = 0
    for st in plan.steps:
        nbytes += 8 * (st.out.size + sum(t.size for t, _ in st.operands if t.kind == 1))
    return nbytes / PlanBuilder.COST_BYTES_PER_NS + loads / PlanBuilder.COST_LOADS_PER_NS


def plan_flops(plan: Plan) -> float:
    """Multiply-adds per evidence set the plan-specialised kernel (csrc/pgx_spec.cu) would emit, roughly: per product
    term one instruction per batch-dependent factor (the batch-invariant ones are folded into one immediate), a dozen
    for a division. The objective for plans that will be specialised: that kernel keeps every message in registers,
    so bytes — what plan_cost prices for the table-driven and step kernels — cost it nothing."""
    ev = set(plan.ev_vars)
    total = 0.0
    for st in plan.steps:
        ssz = 1
        for v in st.sum_vars:
            ssz *= plan.card[v]
        n_dyn = sum(1 for t, d in st.operands if not d and (t.kind == 1 or any(v in ev for v in t.vars)))
        n_const = sum(1 for t, d in st.operands if not d) - n_dyn
        per_term = n_dyn if n_const else max(1, n_dyn - 1)
        total += st.out.size * (ssz * max(1, per_term) + (12 if any(d for _, d in st.operands) else 0))
    return total


def compile_jt_plan(
    jt: JTStructure,
    evidence_vars: Sequence[Hashable] = (),
    variables: Optional[Sequence[Hashable]] = None,
    normalize: bool = True,
    emit_beliefs: bool = False,
    distribute: str = "auto",
    reduce_max: bool = False,
    factorized="auto",
    soft_vars: Sequence[Hashable] = (),
    objective: str = "bytes",
) -> Plan:
    """Two-pass message passing on the rooted junction tree for one evidence-variable signature.

      collect    (leaves -> root):  m[i->p] = sum_{C_i \\ S_ip} psi_i[E=e] * prod_{c in ch(i)} m[c->i]
      distribute (root -> leaves):  m[p->c] = sum_{C_p \\ S_pc} psi_p[E=e] * prod_{n in nb(p) \\ c} m[n->p]
                       or, for cliques of degree >= 3 ("divide"):  beta_p = psi_p * prod_n m[n->p],
                       m[p->c] = (sum_{C_p \\ S_pc} beta_p) / m[c->p]   with 0/0 -> 0  — the reference's
                       belief-update rule sigma / mu (ExactInference.py:788-805, DiscreteFactor.py:859-863)
      marginals: P(v, e) from the cheapest sepset or clique holding v, normalised per variable.

    `variables=None` means every unobserved variable of the tree (all-marginals query).
    `emit_beliefs=True` emits the calibrated clique beliefs and sepset beliefs instead (un-normalised,
    like get_clique_beliefs/get_sepset_beliefs, :750-768).

    `factorized`: psi_i enters every product as the LIST of factors assigned to the clique (the CPDs) instead of one
    clique-sized table, so the re-association of PlanBuilder.contract can multiply a 3-variable CPD into a small
    message before anything clique-sized is touched, and a clique without factors costs nothing at all.
    `distribute`: "ss" | "belief" | "divide" as above, "adaptive" = per clique, whichever of Shafer-Shenoy messages
    and materialised belief + marginalisation lattice the cost model prices lower; "auto" compiles the candidate
    strategies and returns the cheapest plan — cheapest in modelled bytes + loads (`objective="bytes"`, the
    table-driven and step kernels) or in multiply-adds (`objective="flops"`, plans that will be specialised)."""
    if (distribute == "auto" or factorized == "auto") and not emit_beliefs:
        # dense potentials + Shafer-Shenoy keeps the workspace smallest (messages only: fits shared memory for
        # alarm-class models); factor lists + per-clique choice wins on the big-clique models
        dense_ok = sum(int(np.prod([jt.card[v] for v in c], dtype=np.int64)) for c in jt.cliques) <= (1 << 22)
        fz = [False, True] if factorized == "auto" else [bool(factorized)]
        if not dense_ok and factorized == "auto":
            fz = [True]
        cands = []
        # a hub clique (pathfinder: 64 neighbours) makes pure Shafer-Shenoy quadratic in its degree: not a candidate
        hub = max(len(x) for x in jt.nb) > 12
        for f in fz:
            for d in (("ss", "belief", "adaptive") if distribute == "auto" else (distribute,)):
                if d == "adaptive" and not f:
                    continue
                if d == "ss" and hub and distribute == "auto":
                    continue
                cands.append(compile_jt_plan(jt, evidence_vars, variables, normalize, False, d, reduce_max, f, soft_vars))
        if objective == "flops":
            # ask the generator itself (host only, milliseconds): it folds evidence-independent messages and
            # batch-invariant factors, which no formula over the step list sees (hepar2: Shafer-Shenoy around its
            # 17-neighbour hub is 13 540 instructions after folding, the belief-update plan 20 041)
            from .specialize import spec_flops

            if hub and distribute == "auto":
                cands += [compile_jt_plan(jt, evidence_vars, variables, normalize, False, "ss", reduce_max, f, soft_vars) for f in fz]
            scored = [(spec_flops(p), i) for i, p in enumerate(cands)]
            ok = [(c, i) for c, i in scored if c is not None]
            return cands[min(ok)[1]] if ok else min(cands, key=plan_flops)
        cost = [plan_cost(p) for p in cands]
        best = min(range(len(cands)), key=lambda i: cost[i])
        # stay with the (dense, Shafer-Shenoy) candidate — smallest workspace, every step on the fused kernel's fast
        # path — unless another one is clearly cheaper
        first = cands[0].meta
        if first["distribute"] == "ss" and not first["factorized"] and cost[0] <= 1.2 * cost[best]:
            return cands[0]
        return cands[best]
    factorized = bool(factorized) if factorized != "auto" else False
    if distribute == "auto":
        distribute = "belief"  # emit_beliefs materialises every belief anyway
    ev = list(evidence_vars)
    evset = set(ev)
    card = jt.card
    b = PlanBuilder(card, ev)
    n = len(jt.cliques)
    free = [tuple(v for v in c if v not in evset) for c in jt.cliques]
    if factorized:
        psi: List[List[Table]] = []
        for i in range(n):
            mine = [b.add_const(sc, vals) for sc, vals in jt.factors[i] if not all(v in evset for v in sc)]
            psi.append(mine)
    else:
        psi = [[b.add_const(jt.cliques[i], jt.potentials[i])] for i in range(n)]
    for v in soft_vars:
        # soft evidence: a per-evidence-set likelihood vector over v (Plan.inputs) joins the potential of the smallest
        # clique holding v — the reference's observed binary child of v (inference/base.py:256-299) after reduction
        if v in evset:
            raise ValueError(f"soft evidence on the observed variable {v}")
        home = [i for i in range(n) if v in jt.cliques[i]]
        if not home:
            raise ValueError(f"soft-evidence variable {v} is not in the junction tree")
        i = min(home, key=lambda i: (int(np.prod([card[u] for u in free[i]], dtype=np.int64)), i))
        psi[i] = psi[i] + [b.add_input([v])]
    ones: Dict[Hashable, Table] = {}

    def covered(ops, need):
        """operands + all-ones vectors for the variables of `need` that no operand mentions (a clique variable whose
        only factors live in other cliques enters this product through broadcasting)"""
        have = set(v for t in ops for v in t.vars)
        extra = []
        for v in need:
            if v not in have:
                t = ones.get(v)
                if t is None or t.tid >= len(b.tables) or b.tables[t.tid] is not t:  # (a rolled-back trial may have made it)
                    ones[v] = b.add_const([v], np.ones(card[v]))
                extra.append(ones[v])
        return list(ops) + extra

    def sep(i, j):
        sj = set(jt.cliques[j])
        return tuple(v for v in free[i] if v in sj)

    def fsize(vars_):
        s = 1
        for v in vars_:
            s *= card[v]
        return s

    up: Dict[int, Table] = {}
    # collect: children before parents
    for i in reversed(jt.pre):
        p = jt.parent[i]
        if p < 0:
            continue
        ops = psi[i] + [up[c] for c in jt.children[i]]
        up[i] = b.contract(covered(ops, sep(i, p)), sep(i, p), level=jt.height[i], reduce_max=reduce_max)
    base_level = max(jt.height) + 1
    down: Dict[int, Table] = {}
    belief: Dict[int, Table] = {}

    def incoming(i, exclude=None):
        ops = [up[c] for c in jt.children[i] if c != exclude]
        if jt.parent[i] >= 0 and jt.parent[i] != exclude:
            ops.append(down[i])
        return ops

    def want_belief(i):
        if emit_beliefs:
            return True
        if distribute == "ss":
            return False
        if distribute == "divide":
            return len(jt.children[i]) >= 1
        return len(jt.nb[i]) >= 3 and len(jt.children[i]) >= 2  # "belief"

    def messages_from_belief(i, lvl):
        belief[i] = b.contract(covered(psi[i] + incoming(i), free[i]), free[i], level=lvl)
        if fsize(free[i]) >= 4096 and len(jt.children[i]) >= 2:
            # big belief, several children: marginalisation lattice. Child sepsets are served largest first, each from
            # the smallest table already summed whose scope contains it (sum_{C \ S_a} = sum_{S_b \ S_a} sum_{C \ S_b}
            # when S_a is inside S_b), so a 2.7 M-entry munin belief is swept once or twice instead of once per
            # child; the sigma / mu division (0/0 -> 0) is applied to the small result.
            done: List[Tuple[Tuple[Hashable, ...], Table]] = []
            union = [v for v in free[i] if any(v in jt.cliques[c] for c in jt.children[i])]
            biggest = max(fsize(sep(i, c)) for c in jt.children[i])
            if biggest < fsize(union) <= fsize(free[i]) // 4:
                # no child sepset contains the others, but together they span only part of the clique: sum the
                # belief down to that union once, then serve every child from the smaller table
                done.append((tuple(union), b.contract([belief[i]], union, level=lvl + 1, reduce_max=reduce_max)))
            for c in sorted(jt.children[i], key=lambda c: -fsize(sep(i, c))):
                s_c = sep(i, c)
                src = belief[i]
                for scope, tab in done:
                    if set(s_c) <= set(scope) and tab.size < src.size:
                        src = tab
                undivided = b.contract([src], s_c, level=lvl + 1, reduce_max=reduce_max)
                done.append((s_c, undivided))
                down[c] = b.contract([undivided], s_c, divisors=[up[c]], level=lvl + 1, reduce_max=reduce_max)
            return
        for c in jt.children[i]:
            down[c] = b.contract([belief[i]], sep(i, c), divisors=[up[c]], level=lvl + 1, reduce_max=reduce_max)

    def messages_shafer_shenoy(i, lvl):
        for c in jt.children[i]:
            ops = psi[i] + incoming(i, exclude=c)
            down[c] = b.contract(covered(ops, sep(i, c)), sep(i, c), level=lvl + 1, reduce_max=reduce_max)

    for i in jt.pre:
        lvl = base_level + 2 * jt.depth[i]
        if distribute == "adaptive" and not emit_beliefs:
            if not jt.children[i]:
                continue
            if len(jt.nb[i]) > 12:  # hub clique: one belief serves every neighbour
                messages_from_belief(i, lvl)
                continue
            mk = b.mark()
            messages_shafer_shenoy(i, lvl)
            cost_ss = b.cost_since(mk)
            b.rollback(mk)
            messages_from_belief(i, lvl)
            if cost_ss <= b.cost_since(mk):
                b.rollback(mk)
                belief.pop(i, None)
                messages_shafer_shenoy(i, lvl)
            continue
        if want_belief(i):
            messages_from_belief(i, lvl)
        else:
            messages_shafer_shenoy(i, lvl)
    final_level = base_level + 2 * (max(jt.depth) + 1)

    if emit_beliefs:
        for i in range(n):
            b.emit(belief[i], False, free[i])
        for i in range(n):
            p = jt.parent[i]
            if p >= 0:
                # sepset belief mu_ip = m[i->p] * m[p->i]
                mu = b.contract([up[i], down[i]], sep(i, p), level=final_level)
                b.emit(mu, False, sep(i, p))
        return b.finalize({"mode": "jt-max-beliefs" if reduce_max else "jt-beliefs", "evidence_vars": tuple(ev), "root": jt.root})

    if variables is None:
        seen = set()
        variables = []
        for c in jt.cliques:
            for v in c:
                if v not in evset and v not in seen:
                    seen.add(v)
                    variables.append(v)
    variables = list(variables)
    holders: Dict[Hashable, List[int]] = {}
    for i, c in enumerate(free):
        for v in c:
            holders.setdefault(v, []).append(i)
    HIER_MIN = 4096  # tables at least this large hand out their single-variable marginals by recursive halving
    result: Dict[Hashable, Table] = {}
    from_clique: Dict[int, List[Hashable]] = {}
    for v in variables:
        if v in evset:
            raise ValueError(f"{v} is observed")
        if v not in holders:
            raise ValueError(f"variable {v} is not in the junction tree")
        best = None  # (cost, kind, clique)
        for i in holders[v]:
            cost = fsize(free[i]) * (1 if i in belief else 1 + len(jt.nb[i]))
            cand = (cost, 0, i)
            if best is None or cand < best:
                best = cand
            p = jt.parent[i]
            if p >= 0 and v in jt.cliques[p]:
                cand = (fsize(sep(i, p)) * 2, 1, i)
                if cand < best:
                    best = cand
        _, kind, i = best
        if kind == 1:
            result[v] = b.contract([up[i], down[i]], [v], level=final_level)
        elif fsize(free[i]) >= HIER_MIN:
            from_clique.setdefault(i, []).append(v)
        elif i in belief:
            result[v] = b.contract([belief[i]], [v], level=final_level)
        else:
            result[v] = b.contract(covered(psi[i] + incoming(i), [v]), [v], level=final_level)

    def halve(table: Table, scope: List[Hashable], needed: List[Hashable]):
        """All single-variable marginals of `needed` out of one big table in ~2 passes over it (each half of the
        scope is summed out once, recursively) instead of one full pass per variable."""
        if not needed:
            return
        if len(needed) == 1 or fsize(scope) < HIER_MIN or len(scope) < 2:
            for v in needed:
                result[v] = b.contract([table], [v], level=final_level)
            return
        total = float(np.sum([np.log(card[v]) for v in scope]))
        acc, cut = 0.0, 1
        for idx, v in enumerate(scope[:-1]):
            acc += float(np.log(card[v]))
            cut = idx + 1
            if acc >= total / 2:
                break
        for part in (scope[:cut], scope[cut:]):
            sub_needed = [v for v in needed if v in part]
            if sub_needed:
                halve(b.contract([table], part, level=final_level), list(part), sub_needed)

    for i, vs in from_clique.items():
        src = belief[i] if i in belief else b.contract(covered(psi[i] + incoming(i), free[i]), free[i], level=final_level)
        halve(src, list(free[i]), vs)
    for v in variables:
        b.emit(result[v], normalize, [v])
    return b.finalize(
        {"mode": "jt", "evidence_vars": tuple(ev), "variables": tuple(variables), "root": jt.root, "n_cliques": n,
         "distribute": distribute, "factorized": factorized}
    )


def compile_jt_mpe_plan(jt: JTStructure, evidence_vars: Sequence[Hashable] = (), soft_vars: Sequence[Hashable] = ()):
    """Most probable explanation over ALL unobserved variables by max-product on the junction tree with back-pointers
    (SURVEY.md §8f rank 1). The reference's map_query maximises the full joint table (ExactInference.py:609-612,
    :1222-1317) — 10^15 entries on alarm; this computes the same argmax in two passes:

      collect (max):  m[i->p] = max over (C_i minus S_ip) of psi_i[E=e] * prod_{c in ch(i)} m[c->i]
      upward beliefs: beta_i = psi_i[E=e] * prod_{c in ch(i)} m[c->i]      (kept in the workspace, one per clique)
      traceback (k_mpe_traceback, root first): the variables of C_i not fixed by an ancestor take the argmax of beta_i
                      restricted to the already fixed ones (first maximum in C-order, like numpy.argmax).

    Returns (plan, trace, columns): the plan has no output segments; `trace` is the int32 descriptor the traceback
    kernel walks — n_cliques | n_columns | per clique in pre-order: work offset lo, hi | n_axes | n_axes x (column,
    cardinality, stride in beta_i, 1 if the axis is assigned here) — and `columns` the variable of each output column."""
    ev = list(evidence_vars)
    evset = set(ev)
    card = jt.card
    b = PlanBuilder(card, ev)
    n = len(jt.cliques)
    free = [tuple(v for v in c if v not in evset) for c in jt.cliques]
    psi: List[List[Table]] = []
    for i in range(n):
        psi.append([b.add_const(sc, vals) for sc, vals in jt.factors[i] if not all(v in evset for v in sc)])
    for v in soft_vars:
        if v in evset:
            raise ValueError(f"soft evidence on the observed variable {v}")
        home = [i for i in range(n) if v in jt.cliques[i]]
        i = min(home, key=lambda i: (int(np.prod([card[u] for u in free[i]], dtype=np.int64)), i))
        psi[i] = psi[i] + [b.add_input([v])]
    ones: Dict[Hashable, Table] = {}

    def covered(ops, need):
        have = set(v for t in ops for v in t.vars)
        extra = []
        for v in need:
            if v not in have:
                if v not in ones:
                    ones[v] = b.add_const([v], np.ones(card[v]))
                extra.append(ones[v])
        return list(ops) + extra

    up: Dict[int, Table] = {}
    beta: Dict[int, Table] = {}
    for i in reversed(jt.pre):
        ops = covered(psi[i] + [up[c] for c in jt.children[i]], free[i])
        if not ops:
            # every variable of the clique is observed and it has no children: a constant (its factors only scale the
            # joint by a number that does not depend on the unobserved variables, irrelevant for the argmax)
            ops = [b.add_const([], np.ones(()), key=("one",))]
        beta[i] = b.contract(ops, free[i], level=2 * jt.height[i], optimize=False, split=False)
        beta[i].last_step = 1 << 60  # read by the traceback kernel after the last step
        p = jt.parent[i]
        if p >= 0:
            sp = set(jt.cliques[p])
            s_ip = tuple(v for v in free[i] if v in sp)
            up[i] = b.contract([beta[i]], s_ip, level=2 * jt.height[i] + 1, reduce_max=True, optimize=False, split=False)
    plan = b.finalize({"mode": "jt-mpe", "evidence_vars": tuple(ev), "root": jt.root, "n_cliques": n,
                       "soft_vars": tuple(soft_vars)})
    columns: List[Hashable] = []
    col_of: Dict[Hashable, int] = {}
    trace = [n, 0]
    for i in jt.pre:
        t = beta[i]
        strides = {}
        acc = 1
        for v in reversed(t.vars):
            strides[v] = acc
            acc *= card[v]
        lo, hi = int(t.offset) & 0xFFFFFFFF, int(t.offset) >> 32
        trace += [lo - (1 << 32) if lo >= (1 << 31) else lo, hi, len(t.vars)]
        for v in t.vars:
            new = v not in col_of
            if new:
                col_of[v] = len(columns)
                columns.append(v)
            trace += [col_of[v], card[v], strides[v], 1 if new else 0]
    trace[1] = len(columns)
    return plan, np.asarray(trace, dtype=np.int32), columns


# ---------------------------------------------------------------------------------------------
# evidence state mapping (bit-exact with pgmpy/utils/state_name.py:71-84)
# ---------------------------------------------------------------------------------------------
def evidence_to_states(states: Dict[Hashable, list], ev_vars: Sequence[Hashable], evidence_rows) -> np.ndarray:
    """List of {var: state_name} dicts -> int32 [B, k] state indices in `ev_vars` slot order.
    Unknown state names raise KeyError like the reference's greedy path."""
    maps = [{name: i for i, name in enumerate(states[v])} for v in ev_vars]
    out = np.empty((len(evidence_rows), len(ev_vars)), dtype=np.int32)
    for r, row in enumerate(evidence_rows):
        if set(row) != set(ev_vars):
            raise ValueError("every evidence set of a batch must observe exactly the plan's evidence variables")
        for j, v in enumerate(ev_vars):
            out[r, j] = maps[j][row[v]]
    return out
