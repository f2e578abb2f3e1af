"""Host-side factor containers: `DiscreteFactor` and `TabularCPD`.

They mirror the reference's public layout so results are drop-in:
  * values are a dense C-order ndarray shaped by `cardinality`, first variable slowest
    (reference: pgmpy/factors/discrete/DiscreteFactor.py:91-127),
  * `TabularCPD` keeps the child variable first, then the parents (pgmpy/factors/discrete/CPD.py:117-183),
  * state-name <-> index maps follow pgmpy/utils/state_name.py:8-84.

These classes carry data between the user, the planner and the CUDA engine. The factor algebra
(`product`, `marginalize`, `reduce`, `divide`, `normalize`) is NOT computed here: each call builds a
one-step contraction plan and runs it on the GPU through `pgmpy_b200.factor_ops` (no CPU fallback).
"""
from __future__ import annotations

import numbers
from typing import Hashable, Iterable

import numpy as np


def _store_state_names(variables, cardinality, state_names):
    """state_name.py:8-60 semantics: default names are range(card); given names must be unique and
    match the cardinality."""
    if state_names:
        for var, card in zip(variables, cardinality):
            if var not in state_names:
                raise ValueError(f"state names not given for variable {var}")
            names = state_names[var]
            if len(names) != int(card):
                raise ValueError(f"Number of state names must be equal to the cardinality ({var})")
            if len(set(names)) != len(names):
                raise ValueError(f"Repeated statenames for variable: {var}")
        sn = {var: list(state_names[var]) for var in variables}
    else:
        sn = {var: list(range(int(card))) for var, card in zip(variables, cardinality)}
    name_to_no = {var: {name: i for i, name in enumerate(sn[var])} for var in variables}
    no_to_name = {var: {i: name for i, name in enumerate(sn[var])} for var in variables}
    return sn, name_to_no, no_to_name


class DiscreteFactor:
    """Dense table factor phi(variables). See module docstring for layout."""

    def __init__(self, variables, cardinality, values, state_names={}):
        if isinstance(variables, str):
            raise TypeError("Variables: Expected type list or array like, got string")
        values = np.array(values, dtype=np.float64)
        if len(cardinality) != len(variables):
            raise ValueError("Number of elements in cardinality must be equal to number of variables")
        if values.size != int(np.prod(cardinality, dtype=np.int64)):
            raise ValueError(f"Values array must be of size: {int(np.prod(cardinality))}")
        if len(set(variables)) != len(variables):
            raise ValueError("Variable names cannot be same")
        if not isinstance(state_names, dict):
            raise ValueError(f"state_names must be of type dict. Got {type(state_names)}.")
        self.variables = list(variables)
        self.cardinality = np.array(cardinality, dtype=int)
        self.values = values.reshape(tuple(int(c) for c in self.cardinality))
        self.state_names, self.name_to_no, self.no_to_name = _store_state_names(
            self.variables, self.cardinality, state_names
        )

    # ---- metadata (host) -------------------------------------------------------------------
    def scope(self):
        return self.variables

    def get_cardinality(self, variables):
        if isinstance(variables, str):
            raise TypeError("variables: Expected type list or array-like, got type str")
        if not all(var in self.variables for var in variables):
            raise ValueError("Variable not in scope")
        return {var: int(self.cardinality[self.variables.index(var)]) for var in variables}

    def get_state_no(self, var, state_name):
        """state_name.py:71-84 — KeyError on unknown names."""
        return self.name_to_no[var][state_name]

    def get_state_names(self, var, state_no):
        return self.no_to_name[var][state_no]

    def get_value(self, **kwargs):
        """DiscreteFactor.py:182-225 — value for a full assignment given as var=state_name."""
        for var in kwargs:
            if var not in self.variables:
                raise ValueError(f"Variable: {var} doesn't exist in the factor object.")
        index = tuple(self.name_to_no[var][kwargs[var]] for var in self.variables)
        return self.values[index]

    def assignment(self, index):
        """DiscreteFactor.py:268-312 — flat indices -> [(var, state_name), ...]."""
        index = np.array(index)
        max_possible = int(np.prod(self.cardinality)) - 1
        if not all(i <= max_possible for i in index):
            raise IndexError("Index greater than max possible index")
        assignments = np.zeros((len(index), len(self.variables)), dtype=int)
        rev_card = self.cardinality[::-1]
        for i, card in enumerate(rev_card):
            assignments[:, i] = index % card
            index = index // card
        assignments = assignments[:, ::-1]
        return [
            [(var, self.no_to_name[var][int(a)]) for var, a in zip(self.variables, row)] for row in assignments
        ]

    def copy(self):
        out = DiscreteFactor.__new__(DiscreteFactor)
        out.variables = list(self.variables)
        out.cardinality = self.cardinality.copy()
        out.values = self.values.copy()
        out.state_names = {k: list(v) for k, v in self.state_names.items()}
        out.name_to_no = {k: dict(v) for k, v in self.name_to_no.items()}
        out.no_to_name = {k: dict(v) for k, v in self.no_to_name.items()}
        return out

    # ---- algebra: executed by the CUDA engine (pgmpy_b200.factor_ops) -----------------------
    def marginalize(self, variables, inplace=True):
        from . import factor_ops

        return factor_ops.marginalize(self, variables, inplace)

    def maximize(self, variables, inplace=True):
        from . import factor_ops

        return factor_ops.maximize(self, variables, inplace)

    def normalize(self, inplace=True):
        from . import factor_ops

        return factor_ops.normalize(self, inplace)

    def reduce(self, values, inplace=True, show_warnings=True):
        from . import factor_ops

        return factor_ops.reduce(self, values, inplace, show_warnings)

    def product(self, phi1, inplace=True):
        from . import factor_ops

        return factor_ops.product(self, phi1, inplace)

    def divide(self, phi1, inplace=True):
        from . import factor_ops

        return factor_ops.divide(self, phi1, inplace)

    def sum(self, phi1, inplace=True):
        from . import factor_ops

        return factor_ops.add(self, phi1, inplace)

    def __mul__(self, other):
        return self.product(other, inplace=False)

    __rmul__ = __mul__

    def __truediv__(self, other):
        return self.divide(other, inplace=False)

    def __add__(self, other):
        return self.sum(other, inplace=False)

    # ---- comparison (host; mirrors DiscreteFactor.py:1033-1084) -----------------------------
    def __eq__(self, other, atol=1e-08):
        if not (isinstance(self, DiscreteFactor) and isinstance(other, DiscreteFactor)):
            return False
        if set(self.scope()) != set(other.scope()):
            return False
        perm = [other.variables.index(v) for v in self.variables]
        vals = np.transpose(other.values, perm)
        other_card = other.cardinality[perm]
        for axis, var in enumerate(self.variables):
            if set(self.state_names[var]) != set(other.state_names[var]):
                return False
            if self.state_names[var] != other.state_names[var]:
                ref_index = [other.state_names[var].index(s) for s in self.state_names[var]]
                vals = np.take(vals, ref_index, axis=axis)
        if vals.shape != self.values.shape:
            return False
        if not np.allclose(vals, self.values, atol=atol):
            return False
        return bool(np.all(self.cardinality == other_card))

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return id(self)

    def __repr__(self):
        var_card = ", ".join(f"{var}:{card}" for var, card in zip(self.variables, self.cardinality))
        return f"<DiscreteFactor representing phi({var_card}) at {hex(id(self))}>"


class TabularCPD(DiscreteFactor):
    """P(variable | evidence) as a 2-D column-stochastic table; stored like the reference with the
    child axis first (pgmpy/factors/discrete/CPD.py:117-183)."""

    def __init__(self, variable, variable_card, values, evidence=None, evidence_card=None, state_names={}):
        self.variable = variable
        if not isinstance(variable_card, numbers.Integral):
            raise TypeError("Event cardinality must be an integer")
        self.variable_card = int(variable_card)
        variables = [variable]
        cardinality = [self.variable_card]
        if evidence_card is not None:
            if isinstance(evidence_card, numbers.Real):
                raise TypeError("Evidence card must be a list of numbers")
            cardinality.extend(int(c) for c in evidence_card)
        if evidence is not None:
            if isinstance(evidence, str):
                raise TypeError("Evidence must be list, tuple or array of strings.")
            if evidence_card is None:
                raise ValueError("Evidence card must be provided if Evidence is provided!")
            variables.extend(evidence)
            if len(evidence_card) != len(evidence):
                raise ValueError("Length of evidence_card doesn't match length of evidence")
        values = np.array(values, dtype=np.float64)
        if values.ndim != 2:
            raise TypeError("Values must be a 2D list/array")
        expected = (self.variable_card, 1 if evidence is None else int(np.prod(evidence_card)))
        if values.shape != expected:
            raise ValueError(f"values must be of shape {expected}. Got shape: {values.shape}")
        if not isinstance(state_names, dict):
            raise ValueError(f"state_names must be of type dict. Got {type(state_names)}")
        super().__init__(variables, cardinality, values.flatten(), state_names=state_names)

    def get_values(self):
        """2-D view [card(variable), prod(card(parents))] (CPD.py:198-221)."""
        if self.variable in self.variables:
            return self.values.reshape(int(self.cardinality[0]), int(np.prod(self.cardinality[1:])))
        return self.values.reshape(int(np.prod(self.cardinality)), 1)

    def get_evidence(self):
        return self.variables[:0:-1]

    def to_factor(self):
        return DiscreteFactor(self.variables, self.cardinality, self.values, self.state_names)

    def copy(self):
        ev = self.variables[1:] if len(self.variables) > 1 else None
        ev_card = [int(c) for c in self.cardinality[1:]] if len(self.variables) > 1 else None
        return TabularCPD(
            self.variable, self.variable_card, self.get_values(), ev, ev_card, state_names=dict(self.state_names)
        )

    def is_valid_cpd(self):
        """Columns sum to 1 within atol=0.01 (DiscreteFactor.py:955-965)."""
        return bool(np.allclose(self.get_values().sum(axis=0), 1.0, atol=0.01))

    def __repr__(self):
        ev = ", ".join(f"{v}:{c}" for v, c in zip(self.variables[1:], self.cardinality[1:]))
        return f"<TabularCPD representing P({self.variable}:{self.variable_card}{' | ' + ev if ev else ''}) at {hex(id(self))}>"


def as_factor_tuple(f):
    """(variables, cardinality, values, state_names) from our classes or duck-typed pgmpy objects."""
    values = f.values
    if hasattr(values, "detach"):
        values = values.detach().cpu().numpy()
    return (
        list(f.variables),
        [int(c) for c in f.cardinality],
        np.ascontiguousarray(np.asarray(values, dtype=np.float64)),
        {v: list(f.state_names[v]) for v in f.variables},
    )
