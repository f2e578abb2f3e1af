"""Native BIF reader: the on-disk format the bnlearn models come in.

Restates the behaviour of pgmpy/readwrite/BIF.py:34-419 with plain regular expressions (pyparsing, which the
reference needs, is not a dependency here, and the reference's reader takes 2-28 s per model):
  * variables in file order, states as strings (BIF.py:361-390);
  * one probability block per variable; parents in declaration order; a block holding a `table` (or `default`)
    line is read as a flat list reshaped to (card(child), -1) (BIF.py:286-294), otherwise one row of
    card(child) numbers per parent configuration, placed by itertools.product order of the parents' states
    (BIF.py:295-307);
  * CPDs are attached in sorted variable-name order (BIF.py:384), edges parent -> child in block order.
"""
from __future__ import annotations

import gzip
import itertools
import re
from typing import Dict, List

import numpy as np

from .factors import TabularCPD
from .models import DiscreteBayesianNetwork

_NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)"
_NAME = r"[^\s,;(){}|\[\]]+"


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    return re.sub(r"//[^\n]*", " ", text)


class BIFReader:
    def __init__(self, path: str = None, string: str = None):
        if path is not None:
            opener = gzip.open if str(path).endswith(".gz") else open
            with opener(path, "rt") as f:
                text = f.read()
        elif string is not None:
            text = string
        else:
            raise ValueError("Must specify either path or string")
        text = _strip_comments(text)
        m = re.search(r"network\s+(" + _NAME + r"|\"[^\"]*\")", text)
        self.network_name = m.group(1).strip('"') if m else "unknown"
        self.variable_names: List[str] = []
        self.variable_states: Dict[str, List[str]] = {}
        for vm in re.finditer(r"variable\s+(" + _NAME + r")\s*\{(.*?)\}\s*(?=variable|probability|$)", text, flags=re.S):
            name, body = vm.group(1), vm.group(2)
            tm = re.search(r"type\s+discrete\s*\[\s*(\d+)\s*\]\s*\{([^}]*)\}", body, flags=re.S)
            if not tm:
                raise ValueError(f"variable {name}: only discrete variables are supported")
            states = [s.strip().strip('"') for s in tm.group(2).split(",") if s.strip()]
            if len(states) != int(tm.group(1)):
                raise ValueError(f"variable {name}: declared {tm.group(1)} states, found {len(states)}")
            self.variable_names.append(name)
            self.variable_states[name] = states
        self.variable_parents: Dict[str, List[str]] = {}
        self.variable_cpds: Dict[str, np.ndarray] = {}
        self.variable_edges = []
        for pm in re.finditer(r"probability\s*\(([^)]*)\)\s*\{(.*?)\}", text, flags=re.S):
            head, body = pm.group(1), pm.group(2)
            parts = head.split("|")
            child = parts[0].strip()
            parents = [p.strip() for p in parts[1].split(",")] if len(parts) > 1 and parts[1].strip() else []
            card = len(self.variable_states[child])
            statements = [st.strip() for st in body.split(";") if st.strip()]
            if any(re.match(r"(table|default)\b", st) for st in statements):
                vals = [float(x) for st in statements for x in re.findall(_NUM, re.sub(r"\([^)]*\)", " ", st))]
                arr = np.array(vals, dtype=np.float64).reshape(card, -1)
            else:
                n_cols = int(np.prod([len(self.variable_states[p]) for p in parents])) if parents else 1
                arr = np.zeros((card, n_cols))
                rows = {}
                for st in statements:
                    rm = re.match(r"\(([^)]*)\)\s*(.*)", st, flags=re.S)
                    if not rm:
                        raise ValueError(f"probability block of {child}: cannot read '{st[:40]}'")
                    key = tuple(s.strip().strip('"') for s in rm.group(1).split(","))
                    rows[key] = [float(x) for x in re.findall(_NUM, rm.group(2))]
                for col, combo in enumerate(itertools.product(*[self.variable_states[p] for p in parents])):
                    arr[:, col] = rows[combo]
            self.variable_parents[child] = parents
            self.variable_cpds[child] = arr
            self.variable_edges.extend((p, child) for p in parents)

    def get_variables(self):
        return list(self.variable_names)

    def get_states(self):
        return {v: list(s) for v, s in self.variable_states.items()}

    def get_parents(self):
        return {v: list(p) for v, p in self.variable_parents.items()}

    def get_edges(self):
        return [list(e) for e in self.variable_edges]

    def get_values(self):
        return dict(self.variable_cpds)

    def get_model(self, state_name_type=str) -> DiscreteBayesianNetwork:
        model = DiscreteBayesianNetwork()
        model.add_nodes_from(self.variable_names)
        for p, c in self.variable_edges:
            model._parents[c].append(p)
            model._children[p].append(c)
        model.name = self.network_name
        for var in sorted(self.variable_cpds):
            parents = self.variable_parents[var]
            sn = {p: [state_name_type(s) for s in self.variable_states[p]] for p in parents}
            sn[var] = [state_name_type(s) for s in self.variable_states[var]]
            model.add_cpds(
                TabularCPD(
                    var, len(self.variable_states[var]), self.variable_cpds[var], parents or None,
                    [len(self.variable_states[p]) for p in parents] or None, state_names=sn,
                )
            )
        return model
