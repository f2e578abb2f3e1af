"""Host graph algorithms used by the planner (pure Python, deterministic, no networkx).

Restated behaviour (not code) of the reference's graph layer:
  * `active_trail_nodes`  — Koller & Friedman Alg. 3.1 as in pgmpy/base/DAG.py:864-950,
  * `ancestors_of`        — pgmpy/base/DAG.py:952-985,
  * `prune_nodes`         — the kept-node set of Inference._prune_bayesian_model, pgmpy/inference/base.py:184-197,
  * `moral_graph`         — pgmpy/base/DAG.py:449-476.
Ours (the reference's own versions are unusable as plan sources):
  * `min_fill_order`      — greedy min-fill WITH fill-in edges, ties by min weight then position,
  * `junction_tree`       — elimination cliques -> maximal cliques -> max-|sepset| spanning tree.
"""
from __future__ import annotations

import heapq
from typing import Dict, Hashable, Iterable, List, Sequence, Set, Tuple


def ancestors_of(parents: Dict[Hashable, Sequence[Hashable]], nodes: Iterable[Hashable]) -> Set[Hashable]:
    """All ancestors of `nodes`, including the nodes themselves."""
    seen = set()
    stack = list(nodes)
    for n in stack:
        if n not in parents:
            raise ValueError(f"Node {n} not in graph")
    while stack:
        n = stack.pop()
        if n in seen:
            continue
        seen.add(n)
        stack.extend(p for p in parents[n] if p not in seen)
    return seen


def active_trail_nodes(parents, children, start, observed: Set[Hashable]) -> Set[Hashable]:
    """Nodes reachable from `start` by an active trail given `observed` (start included unless
    observed). Direction labels follow the reference: "up" = arrived from a child."""
    anc = ancestors_of(parents, observed) if observed else set()
    visit = [(start, "up")]
    traversed = set()
    active = set()
    while visit:
        node, direction = visit.pop()
        if (node, direction) in traversed:
            continue
        traversed.add((node, direction))
        if node not in observed:
            active.add(node)
        if direction == "up" and node not in observed:
            for p in parents[node]:
                visit.append((p, "up"))
            for c in children[node]:
                visit.append((c, "down"))
        elif direction == "down":
            if node not in observed:
                for c in children[node]:
                    visit.append((c, "down"))
            if node in anc:
                for p in parents[node]:
                    visit.append((p, "up"))
    return active


def prune_nodes(parents, children, variables: Sequence[Hashable], evidence_vars: Sequence[Hashable]) -> Set[Hashable]:
    """Kept-node set K of the reference's query pruning:
    S = U_q active_trail_nodes(q | E)  U  E;  H = subgraph(S);  K = ancestors within H of (Q U E)."""
    observed = set(evidence_vars)
    s = set(observed)
    for q in variables:
        s |= active_trail_nodes(parents, children, q, observed)
    sub_parents = {n: [p for p in parents[n] if p in s] for n in s}
    targets = [v for v in list(variables) + list(evidence_vars) if v in s]
    return ancestors_of(sub_parents, targets)


def moral_graph(parents: Dict[Hashable, Sequence[Hashable]]) -> Dict[Hashable, Set[Hashable]]:
    adj: Dict[Hashable, Set[Hashable]] = {n: set() for n in parents}
    for n, ps in parents.items():
        ps = list(ps)
        for p in ps:
            adj[n].add(p)
            adj[p].add(n)
        for i in range(len(ps)):
            for j in range(i + 1, len(ps)):
                adj[ps[i]].add(ps[j])
                adj[ps[j]].add(ps[i])
    return adj


def interaction_graph(scopes: Iterable[Sequence[Hashable]]) -> Dict[Hashable, Set[Hashable]]:
    adj: Dict[Hashable, Set[Hashable]] = {}
    for sc in scopes:
        sc = list(sc)
        for v in sc:
            adj.setdefault(v, set())
        for i in range(len(sc)):
            for j in range(i + 1, len(sc)):
                adj[sc[i]].add(sc[j])
                adj[sc[j]].add(sc[i])
    return adj


def _fill_count(adj, v):
    nb = list(adj[v])
    cnt = 0
    for i in range(len(nb)):
        a = adj[nb[i]]
        for j in range(i + 1, len(nb)):
            if nb[j] not in a:
                cnt += 1
    return cnt


def min_fill_order(
    adj: Dict[Hashable, Set[Hashable]],
    card: Dict[Hashable, int],
    keep: Iterable[Hashable] = (),
    rank: Dict[Hashable, int] | None = None,
) -> Tuple[List[Hashable], List[Tuple[Hashable, ...]]]:
    """Greedy min-fill elimination of every node not in `keep`.

    Cost = (#fill-in edges, weight of the created clique, position) — fully deterministic.
    Returns (order, elimination cliques) where clique[i] = (order[i],) + its neighbours at elimination
    time. The graph is copied; fill-in edges ARE added (unlike the reference's MinFill,
    pgmpy/inference/EliminationOrder.py:107-116,160-166)."""
    adj = {v: set(n) for v, n in adj.items()}
    keep = set(keep)
    if rank is None:
        rank = {v: i for i, v in enumerate(adj)}

    def cost(v):
        w = card[v]
        for n in adj[v]:
            w *= card[n]
        return (_fill_count(adj, v), w, rank[v])

    heap = [(cost(v), v) for v in adj if v not in keep]
    heapq.heapify(heap)
    current = {v: c for c, v in heap}
    order, cliques = [], []
    alive = set(adj)
    while heap:
        c, v = heapq.heappop(heap)
        if v not in alive or current.get(v) != c:
            continue
        nb = sorted(adj[v], key=lambda x: rank[x])
        order.append(v)
        cliques.append((v,) + tuple(nb))
        touched = set(nb)
        for i in range(len(nb)):
            for j in range(i + 1, len(nb)):
                if nb[j] not in adj[nb[i]]:
                    adj[nb[i]].add(nb[j])
                    adj[nb[j]].add(nb[i])
        for n in nb:
            adj[n].discard(v)
            touched |= adj[n]
        alive.discard(v)
        del adj[v]
        current.pop(v, None)
        for n in touched:
            if n in alive and n not in keep:
                nc = cost(n)
                if current.get(n) != nc:
                    current[n] = nc
                    heapq.heappush(heap, (nc, n))
    return order, cliques


def junction_tree(
    adj: Dict[Hashable, Set[Hashable]], card: Dict[Hashable, int], rank: Dict[Hashable, int] | None = None
) -> Tuple[List[Tuple[Hashable, ...]], List[Tuple[int, int]]]:
    """Junction tree of an undirected graph via min-fill triangulation.

    Returns (cliques, edges): cliques are tuples of variables (the maximal elimination cliques,
    members ordered by `rank`), edges are index pairs forming ONE tree (components of a disconnected
    graph are linked by empty sepsets).

    Construction: elimination clique C_i = {v_i} + neighbours of v_i when eliminated; its tree parent
    is the clique of the earliest-eliminated member of C_i - {v_i}; this elimination tree has the
    running-intersection property. A parent contained in one of its (possibly inherited) children is
    not maximal and is merged into that child (the child inherits the parent's other links)."""
    if rank is None:
        rank = {v: i for i, v in enumerate(adj)}
    order, elim = min_fill_order(adj, card, keep=(), rank=rank)
    pos = {v: i for i, v in enumerate(order)}
    n = len(order)
    sets = [frozenset(c) for c in elim]
    parent = [-1] * n
    kids: List[List[int]] = [[] for _ in range(n)]
    for i, cl in enumerate(elim):
        if len(cl) > 1:
            parent[i] = min(pos[v] for v in cl[1:])
            kids[parent[i]].append(i)
    merged_into = [-1] * n
    # parents always have a larger elimination index than their children: go top-down
    for p in range(n - 1, -1, -1):
        target = -1
        for c in kids[p]:
            if sets[p] <= sets[c]:
                target = c
                break
        if target < 0:
            continue
        merged_into[p] = target
        parent[target] = parent[p]
        if parent[p] >= 0:
            gp = parent[p]
            kids[gp] = [target if k == p else k for k in kids[gp]]
        for c in kids[p]:
            if c != target:
                parent[c] = target
                kids[target].append(c)
        kids[p] = []
    roots = [i for i in range(n) if merged_into[i] < 0]
    index = {r: k for k, r in enumerate(roots)}
    cliques = [tuple(sorted(elim[r], key=lambda x: rank[x])) for r in roots]
    edges = sorted(
        (min(index[i], index[parent[i]]), max(index[i], index[parent[i]])) for i in roots if parent[i] >= 0
    )
    # join the components of a forest with empty-sepset edges
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
        edges.append((min(reps[0], reps[k]), max(reps[0], reps[k])))
    return cliques, sorted(edges)


def check_running_intersection(cliques, edges) -> bool:
    """True iff for every variable the cliques containing it form a connected subtree."""
    nb = {i: set() for i in range(len(cliques))}
    for a, b in edges:
        nb[a].add(b)
        nb[b].add(a)
    if len(edges) != len(cliques) - 1:
        return False
    allvars = set()
    for c in cliques:
        allvars |= set(c)
    for v in allvars:
        holders = {i for i, c in enumerate(cliques) if v in c}
        start = next(iter(holders))
        seen = {start}
        stack = [start]
        while stack:
            x = stack.pop()
            for y in nb[x]:
                if y in holders and y not in seen:
                    seen.add(y)
                    stack.append(y)
        if seen != holders:
            return False
    return True
