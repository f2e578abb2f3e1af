"""Stand-in for opt_einsum (not installed, not pinned by the reference: `opt_einsum>=3.3`,
/root/reference/pyproject.toml:36). The reference only uses `contract(..., optimize="greedy")`,
which picks a pairwise order and calls numpy; results differ between orders at ~1e-16 relative.
Here: map hashable labels to ints and defer to np.einsum(optimize="greedy")."""
import numpy as np


def _to_int_labels(args):
    ops = list(args)
    out = None
    if len(ops) % 2 == 1:
        out = ops.pop()
    table = {}
    new = []
    for i in range(0, len(ops), 2):
        labels = [table.setdefault(l, len(table)) for l in ops[i + 1]]
        new.extend([ops[i], labels])
    if out is not None:
        new.append([table.setdefault(l, len(table)) for l in out])
    return new


def contract(*operands, optimize="greedy", **kwargs):
    if isinstance(operands[0], str):
        return np.einsum(*operands, optimize="greedy")
    new = _to_int_labels(operands)
    try:
        import torch

        if any(isinstance(o, torch.Tensor) for o in new[0::2]):
            return torch.einsum(*new)
    except ImportError:  # pragma: no cover
        pass
    return np.einsum(*new, optimize="greedy")


def contract_path(*operands, **kwargs):
    if isinstance(operands[0], str):
        return np.einsum_path(*operands, optimize="greedy")
    return np.einsum_path(*_to_int_labels(operands), optimize="greedy")
