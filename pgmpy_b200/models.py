"""Host-side model containers: `DiscreteBayesianNetwork` and `JunctionTree`.

Only what the exact-inference path touches is mirrored (reference: pgmpy/models/DiscreteBayesianNetwork.py
add_cpds:244, get_cpds:305, get_cardinality:387, states:436, check_model:451, to_junction_tree:539;
pgmpy/models/JunctionTree.py, ClusterGraph.py add_factors:130, get_factors:166, check_model:329).
Learning, sampling, IO and causal APIs of the reference are out of scope.

`from_pgmpy` adapts a duck-typed pgmpy model so the engine can be used directly on reference objects.
"""
from __future__ import annotations

import json
import os
from typing import Dict, Hashable, List, Sequence, Tuple

import numpy as np

from . import graph as G
from .factors import DiscreteFactor, TabularCPD, as_factor_tuple


class DiscreteBayesianNetwork:
    def __init__(self, ebunch=None):
        self._nodes: List[Hashable] = []
        self._parents: Dict[Hashable, List[Hashable]] = {}
        self._children: Dict[Hashable, List[Hashable]] = {}
        self.cpds: List[TabularCPD] = []
        if ebunch:
            self.add_edges_from(ebunch)

    # ---- graph -----------------------------------------------------------------------------
    def add_node(self, node):
        if node not in self._parents:
            self._nodes.append(node)
            self._parents[node] = []
            self._children[node] = []

    def add_nodes_from(self, nodes):
        for n in nodes:
            self.add_node(n)

    def add_edge(self, u, v):
        if u == v:
            raise ValueError("Self loops are not allowed.")
        self.add_node(u)
        self.add_node(v)
        if v in G.ancestors_of(self._parents, [u]):
            raise ValueError(f"Loops are not allowed. Adding the edge from ({u}->{v}) forms a loop.")
        if u not in self._parents[v]:
            self._parents[v].append(u)
            self._children[u].append(v)

    def add_edges_from(self, ebunch):
        for u, v in ebunch:
            self.add_edge(u, v)

    def nodes(self):
        return list(self._nodes)

    def edges(self):
        return [(p, c) for c in self._nodes for p in self._parents[c]]

    def get_parents(self, node):
        return list(self._parents[node])

    def get_children(self, node):
        return list(self._children[node])

    def predecessors(self, node):
        return iter(self._parents[node])

    def successors(self, node):
        return iter(self._children[node])

    def __contains__(self, node):
        return node in self._parents

    def __len__(self):
        return len(self._nodes)

    # ---- CPDs ------------------------------------------------------------------------------
    def add_cpds(self, *cpds):
        for cpd in cpds:
            if not isinstance(cpd, TabularCPD):
                raise ValueError("Only TabularCPD can be added.")
            if set(cpd.scope()) - set(self._nodes):
                raise ValueError("CPD defined on variable not in the model", cpd)
            for i, prev in enumerate(self.cpds):
                if prev.variable == cpd.variable:
                    self.cpds[i] = cpd
                    break
            else:
                self.cpds.append(cpd)

    def get_cpds(self, node=None):
        if node is None:
            return list(self.cpds)
        if node not in self._parents:
            raise ValueError("Node not present in the Directed Graph")
        for cpd in self.cpds:
            if cpd.variable == node:
                return cpd
        return None

    def remove_cpds(self, *cpds):
        for cpd in cpds:
            if isinstance(cpd, (str, int)):
                cpd = self.get_cpds(cpd)
            self.cpds.remove(cpd)

    def get_cardinality(self, node=None):
        if node is not None:
            return self.get_cpds(node).variable_card
        return {cpd.variable: cpd.variable_card for cpd in self.cpds}

    @property
    def states(self):
        out = {}
        for cpd in self.cpds:
            out.update({v: list(s) for v, s in cpd.state_names.items()})
        return out

    def check_model(self):
        """DiscreteBayesianNetwork.py:451-508: every node has a CPD over exactly its family, columns sum
        to 1 (atol 0.01), parent cardinalities and state names agree with the parents' own CPDs."""
        for node in self._nodes:
            cpd = self.get_cpds(node)
            if cpd is None:
                raise ValueError(f"No CPD associated with {node}")
            if set(cpd.variables[1:]) != set(self._parents[node]):
                raise ValueError(f"CPD associated with {node} doesn't have proper parents associated with it.")
            if not cpd.is_valid_cpd():
                raise ValueError(f"Sum or integral of conditional probabilities for node {node} is not equal to 1.")
        for node in self._nodes:
            cpd = self.get_cpds(node)
            for idx, parent in enumerate(cpd.variables[1:], start=1):
                pc = self.get_cpds(parent)
                if int(cpd.cardinality[idx]) != pc.variable_card:
                    raise ValueError(f"The cardinality of {parent} doesn't match in it's child nodes.")
                if cpd.state_names[parent] != pc.state_names[parent]:
                    raise ValueError(f"The state names of {parent} doesn't match in it's child nodes.")
        return True

    def copy(self):
        m = DiscreteBayesianNetwork()
        m.add_nodes_from(self._nodes)
        m.add_edges_from(self.edges())
        m.cpds = [c.copy() for c in self.cpds]
        return m

    # ---- batch callers: the reference loops over rows, we bucket by signature ----------
    def _check_prediction_frame(self, data):
        if set(data.columns) == set(self._nodes):
            raise ValueError("No variable missing in data. Nothing to predict")
        if set(data.columns) - set(self._nodes):
            raise ValueError("Data has variables which are not in the model")
        return [v for v in self._nodes if v not in set(data.columns)]

    def predict(self, data, algo=None, stochastic=False, n_jobs=-1, seed=None, **kwargs):
        """States with the highest posterior probability (MAP over the joint of the missing variables) for every
        row of `data` — pgmpy/models/DiscreteBayesianNetwork.py:731-910. Rows are grouped by which columns are
        observed (NaN = missing for that row) and every group is ONE batched map_query on the GPU instead of one
        query per unique row. `stochastic=True` (sampling from the posterior) is outside the accelerated path."""
        import pandas as pd

        from .inference import VariableElimination

        if stochastic:
            raise NotImplementedError("stochastic prediction samples from the posterior; only MAP prediction is accelerated")
        missing = self._check_prediction_frame(data)
        infer = (algo or VariableElimination)(self)
        states = self.states
        out = pd.DataFrame(index=data.index, columns=list(data.columns) + missing, dtype=object)
        for c in data.columns:
            out[c] = data[c]
        observed_mask = data.notna()
        for pattern, idx in observed_mask.groupby(list(data.columns), sort=False).groups.items():
            pattern = pattern if isinstance(pattern, tuple) else (pattern,)
            ev_vars = [c for c, seen in zip(data.columns, pattern) if seen]
            variables = missing + [c for c, seen in zip(data.columns, pattern) if not seen]
            rows = data.loc[idx]
            maps = [{name: i for i, name in enumerate(states[v])} for v in ev_vars]
            ev = np.array([[mp[val] for mp, val in zip(maps, row)] for row in rows[ev_vars].itertuples(index=False)],
                          dtype=np.int32).reshape(len(rows), len(ev_vars))
            pred = infer.map_query_batch(variables, ev_vars, ev).cpu().numpy()
            for j, v in enumerate(variables):
                out.loc[idx, v] = [states[v][int(i)] for i in pred[:, j]]
        return out

    def predict_probability(self, data):
        """Posterior probability of every state of every missing variable, one row per data row, columns
        `<var>_<state>` — pgmpy/models/DiscreteBayesianNetwork.py:912-989 (marginals of the joint posterior over
        the missing variables). One batched query instead of the reference's per-row loop."""
        import pandas as pd

        from .inference import VariableElimination

        missing = self._check_prediction_frame(data)
        infer = VariableElimination(self)
        states = self.states
        ev_vars = list(data.columns)
        maps = [{name: i for i, name in enumerate(states[v])} for v in ev_vars]
        ev = np.array([[mp[val] for mp, val in zip(maps, row)] for row in data.itertuples(index=False)],
                      dtype=np.int32).reshape(len(data), len(ev_vars))
        res = infer.query_batch(missing, ev_vars, ev, joint=False).cpu().numpy()
        cp = infer._plan(missing, ev_vars, False, None)
        cols = {}
        for seg in cp.plan.segments:
            v = seg.vars[0]
            for i, st in enumerate(states[v]):
                cols[f"{v}_{st}"] = res[:, seg.out_offset + i]
        return pd.DataFrame(cols, index=data.index)

    # ---- junction tree (our min-fill builder; the reference's H6 builder is unusable) --
    def to_junction_tree(self) -> "JunctionTree":
        from .planner import build_junction_tree

        return build_junction_tree(self)


class JunctionTree:
    """Clique tree container: nodes are tuples of variables, one potential (DiscreteFactor) per clique."""

    def __init__(self, ebunch=None):
        self._nodes: List[Tuple[Hashable, ...]] = []
        self._adj: Dict[Tuple[Hashable, ...], List[Tuple[Hashable, ...]]] = {}
        self.factors: List[DiscreteFactor] = []
        if ebunch:
            for u, v in ebunch:
                self.add_edge(u, v)

    def add_node(self, node):
        if not isinstance(node, (list, set, tuple)):
            raise TypeError("Node can only be a list, set or tuple of nodes forming a clique")
        node = tuple(node)
        if node not in self._adj:
            self._nodes.append(node)
            self._adj[node] = []
        return node

    def add_nodes_from(self, nodes):
        for n in nodes:
            self.add_node(n)

    def add_edge(self, u, v):
        u, v = self.add_node(u), self.add_node(v)
        if v in self._reachable(u):
            raise ValueError(f"Addition of edge between {u} and {v} forms a cycle breaking the properties of Junction Tree")
        self._adj[u].append(v)
        self._adj[v].append(u)

    def _reachable(self, start):
        seen = {start}
        stack = [start]
        while stack:
            x = stack.pop()
            for y in self._adj[x]:
                if y not in seen:
                    seen.add(y)
                    stack.append(y)
        return seen

    def nodes(self):
        return list(self._nodes)

    def edges(self):
        seen = set()
        out = []
        for u in self._nodes:
            for v in self._adj[u]:
                if (v, u) not in seen:
                    seen.add((u, v))
                    out.append((u, v))
        return out

    def neighbors(self, node):
        return list(self._adj[tuple(node)])

    def add_factors(self, *factors):
        for f in factors:
            scope = set(f.scope())
            if not any(scope == set(n) for n in self._nodes):
                raise ValueError("Factors defined on clusters of variable not present in model")
            self.factors.append(f)

    def get_factors(self, node=None):
        if node is None:
            return list(self.factors)
        for f in self.factors:
            if set(f.scope()) == set(node):
                return f
        return None

    def get_cardinality(self, node=None):
        card = {}
        for f in self.factors:
            for v, c in zip(f.variables, f.cardinality):
                card[v] = int(c)
        return card if node is None else card[node]

    @property
    def states(self):
        out = {}
        for f in self.factors:
            out.update({v: list(s) for v, s in f.state_names.items()})
        return out

    def check_model(self):
        if self._nodes and len(self._reachable(self._nodes[0])) != len(self._nodes):
            raise ValueError("The Junction Tree defined is not fully connected.")
        for n in self._nodes:
            if self.get_factors(n) is None:
                raise ValueError("Factors for all the cliques or clusters not defined.")
        card = {}
        for f in self.factors:
            for v, c in zip(f.variables, f.cardinality):
                if card.setdefault(v, int(c)) != int(c):
                    raise ValueError(f"Cardinality of variable {v} not matching among factors")
        return True

    def copy(self):
        jt = JunctionTree()
        jt.add_nodes_from(self._nodes)
        for u, v in self.edges():
            jt.add_edge(u, v)
        jt.factors = [f.copy() for f in self.factors]
        return jt


# ---- adapters & loaders -----------------------------------------------------------------------
def from_pgmpy(model) -> DiscreteBayesianNetwork:
    """Adapt a (duck-typed) pgmpy DiscreteBayesianNetwork: same node order, same CPD axis order."""
    if isinstance(model, DiscreteBayesianNetwork):
        return model
    m = DiscreteBayesianNetwork()
    m.add_nodes_from(list(model.nodes()))
    for u, v in model.edges():
        m.add_node(u)
        m.add_node(v)
        m._parents[v].append(u)
        m._children[u].append(v)
    for node in model.nodes():
        cpd = model.get_cpds(node)
        variables, card, values, sn = as_factor_tuple(cpd)
        m.cpds.append(
            TabularCPD(
                variables[0], card[0], values.reshape(card[0], -1), variables[1:] or None, card[1:] or None, sn
            )
        )
    return m


def junction_tree_from_pgmpy(jt) -> JunctionTree:
    if isinstance(jt, JunctionTree):
        return jt
    out = JunctionTree()
    out.add_nodes_from(list(jt.nodes()))
    for u, v in jt.edges():
        out.add_edge(u, v)
    for f in jt.get_factors():
        variables, card, values, sn = as_factor_tuple(f)
        out.factors.append(DiscreteFactor(variables, card, values, sn))
    return out


MODEL_DIR = os.environ.get(
    "PGX_MODEL_DIR",
    os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "models"),  # package data
)


def load_model_npz(path: str) -> DiscreteBayesianNetwork:
    """Load a model fixture written by oracle/export_models.py (JSON header + packed fp64 CPT blob)."""
    with np.load(path, allow_pickle=False) as z:
        header = json.loads(str(z["header"]))
        values = z["values"]
    m = DiscreteBayesianNetwork()
    m.add_nodes_from(header["nodes"])
    for u, v in header["edges"]:
        m._parents[v].append(u)
        m._children[u].append(v)
    for c in header["cpds"]:
        card = c["cardinality"]
        vals = values[c["offset"] : c["offset"] + c["size"]].reshape(card[0], -1)
        m.cpds.append(
            TabularCPD(
                c["variable"], card[0], vals, c["variables"][1:] or None, card[1:] or None, state_names=c["state_names"]
            )
        )
    return m


def get_example_model(name: str) -> DiscreteBayesianNetwork:
    """bnlearn example model by name (asia, alarm, hepar2, win95pts, pathfinder, munin, diabetes, ...),
    mirroring pgmpy.utils.get_example_model (pgmpy/utils/utils.py:16) for the shipped fixtures."""
    path = os.path.join(MODEL_DIR, name + ".npz")
    if not os.path.exists(path):
        raise ValueError(f"example model {name!r} not available (looked in {MODEL_DIR})")
    return load_model_npz(path)
