"""pgmpy_b200 — B200-native exact inference for pgmpy-style discrete Bayesian networks.

Drop-in surface for ONE hot path of pgmpy (reference: tristantreb/pgmpy v1.0.0):
`VariableElimination.query`, `BeliefPropagation.calibrate/query`, `DiscreteFactor` tables.
Host Python compiles a model (+ query signature) into a static contraction plan; hand-written
sm_100a CUDA kernels behind a C-ABI (`libpgx.so`, include/pgx.h) execute it over a batch of
independent evidence sets. There is no CPU fallback: without the CUDA library every compute
call raises.
"""
from .config import config
from .factors import DiscreteFactor, TabularCPD
from .models import DiscreteBayesianNetwork, JunctionTree, get_example_model, from_pgmpy

__all__ = [
    "DiscreteFactor",
    "TabularCPD",
    "DiscreteBayesianNetwork",
    "JunctionTree",
    "get_example_model",
    "from_pgmpy",
    "config",
    "VariableElimination",
    "BeliefPropagation",
]
__version__ = "0.1.0"


def __getattr__(name):
    if name in ("VariableElimination", "BeliefPropagation"):
        from . import inference

        return getattr(inference, name)
    raise AttributeError(name)
