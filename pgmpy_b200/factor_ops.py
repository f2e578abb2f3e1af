"""DiscreteFactor algebra on the GPU: each call is a one-step contraction plan run through libpgx.so.

Mirrors the semantics of pgmpy/factors/discrete/DiscreteFactor.py — marginalize :360-411, maximize :413-483,
normalize :485-533, reduce :535-617, product :717-792, divide :794-866 (0/0 -> 0, x/0 -> inf) — including
the inplace convention (inplace=True mutates and returns None). These calls move one table to the device
and back, so they are for API completeness and parity tests; the throughput path is the batched plans of
`pgmpy_b200.inference`.
"""
from __future__ import annotations

import numbers

import numpy as np

from .engine import CompiledPlan
from .plan import PlanBuilder


def _run_single(builder: PlanBuilder, table, normalize=False, ev_states=None) -> np.ndarray:
    builder.emit(table, normalize)
    plan = builder.finalize({"mode": "factor-op"})
    cp = CompiledPlan(plan)
    ev = np.zeros((1, len(plan.ev_vars)), dtype=np.int32) if ev_states is None else np.asarray(ev_states, dtype=np.int32)
    return cp.run_host(ev)[0]


def _card(*factors):
    card = {}
    for f in factors:
        for v, c in zip(f.variables, f.cardinality):
            if card.setdefault(v, int(c)) != int(c):
                raise ValueError(f"cardinality of {v} differs between the factors")
    return card


def _assign(phi, variables, values, state_names):
    phi.variables = list(variables)
    phi.cardinality = np.array([values.shape[i] for i in range(values.ndim)], dtype=int)
    phi.values = values
    phi.state_names = {v: list(state_names[v]) for v in variables}
    phi.name_to_no = {v: {n: i for i, n in enumerate(phi.state_names[v])} for v in variables}
    phi.no_to_name = {v: {i: n for i, n in enumerate(phi.state_names[v])} for v in variables}


def _reduce_axes(phi, variables, inplace, use_max):
    if isinstance(variables, str):
        raise TypeError("variables: Expected type list or array-like, got type str")
    out = phi if inplace else phi.copy()
    for var in variables:
        if var not in out.variables:
            raise ValueError(f"{var} not in scope.")
    keep = [v for v in out.variables if v not in variables]
    card = _card(out)
    b = PlanBuilder(card, [])
    t = b.add_const(out.variables, out.values)
    res = b.contract([t], keep, reduce_max=use_max)
    vals = _run_single(b, res).reshape([card[v] for v in keep])
    _assign(out, keep, vals, out.state_names)
    if not inplace:
        return out


def marginalize(phi, variables, inplace=True):
    return _reduce_axes(phi, variables, inplace, use_max=False)


def maximize(phi, variables, inplace=True):
    return _reduce_axes(phi, variables, inplace, use_max=True)


def normalize(phi, inplace=True):
    out = phi if inplace else phi.copy()
    card = _card(out)
    b = PlanBuilder(card, [])
    t = b.add_const(out.variables, out.values)
    res = b.contract([t], out.variables)
    out.values = _run_single(b, res, normalize=True).reshape(out.values.shape)
    if not inplace:
        return out


def reduce(phi, values, inplace=True, show_warnings=True):
    if isinstance(values, str):
        raise TypeError("values: Expected type list or array-like, got type str")
    if not all(isinstance(t, tuple) for t in values):
        raise TypeError("values: Expected type list of tuples")
    for var, _ in values:
        if var not in phi.variables:
            raise ValueError(f"The variable: {var} is not in the factor")
    out = phi if inplace else phi.copy()
    try:
        pairs = [(var, out.get_state_no(var, s)) for var, s in values]
    except KeyError:
        # unknown names are retried as raw state numbers (DiscreteFactor.py:589-597)
        pairs = [(var, int(s)) for var, s in values]
    ev_vars = [v for v, _ in pairs]
    card = _card(out)
    for v, s in pairs:
        if not -card[v] <= s < card[v]:
            raise IndexError(f"index {s} is out of bounds for variable {v} with {card[v]} states")
    states = [[s % card[v] for v, s in pairs]]  # negative ints wrap like numpy
    keep = [v for v in out.variables if v not in ev_vars]
    b = PlanBuilder(card, ev_vars)
    t = b.add_const(out.variables, out.values)
    res = b.contract([t], keep)
    vals = _run_single(b, res, ev_states=states).reshape([card[v] for v in keep])
    _assign(out, keep, vals, out.state_names)
    if not inplace:
        return out


def product(phi, phi1, inplace=True):
    out = phi if inplace else phi.copy()
    if isinstance(phi1, numbers.Number):
        card = _card(out)
        b = PlanBuilder(card, [])
        res = b.contract([b.add_const(out.variables, out.values), b.add_const((), np.array(float(phi1)))], out.variables)
        out.values = _run_single(b, res).reshape(out.values.shape)
    else:
        card = _card(out, phi1)
        new_vars = list(out.variables) + [v for v in phi1.variables if v not in out.variables]
        b = PlanBuilder(card, [])
        res = b.contract([b.add_const(out.variables, out.values), b.add_const(phi1.variables, phi1.values)], new_vars)
        vals = _run_single(b, res).reshape([card[v] for v in new_vars])
        names = dict(phi1.state_names)
        names.update(out.state_names)
        _assign(out, new_vars, vals, names)
    if not inplace:
        return out


def divide(phi, phi1, inplace=True):
    out = phi if inplace else phi.copy()
    if set(phi1.variables) - set(out.variables):
        raise ValueError("Scope of divisor should be a subset of dividend")
    card = _card(out, phi1)
    b = PlanBuilder(card, [])
    res = b.contract([b.add_const(out.variables, out.values)], out.variables,
                     divisors=[b.add_const(phi1.variables, phi1.values)])
    out.values = _run_single(b, res).reshape(out.values.shape)
    if not inplace:
        return out


def add(phi, phi1, inplace=True):
    raise NotImplementedError("DiscreteFactor.sum is not on the exact-inference path (not used by VE/BP)")
