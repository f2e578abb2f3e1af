"""Global configuration seam: the engine-side counterpart of `pgmpy.config` (pgmpy/global_vars.py:32-189).

The reference selects its array backend through one module-global object: `config.set_backend("numpy" | "torch",
device, dtype)`, `config.set_dtype`, `config.set_show_progress`; factors read `config.get_dtype()` when they are built
(pgmpy/factors/discrete/DiscreteFactor.py:95-102). This module keeps that surface and adds the one backend this package
has, "b200":

    from pgmpy_b200 import config
    config.set_backend("b200", device="cuda:0", dtype="float32")     # same call shape as pgmpy's
    VariableElimination(model)                                        # runs in fp32 (1e-5 mode) on cuda:0

Precedence for the arithmetic type of a new inference object: an explicit `dtype=` argument, then this module's
`config` if the user changed it, then — when the real pgmpy has been imported by the application — whatever
`pgmpy.config.get_dtype()` says (numpy / torch dtypes and their names are all understood), then float64.
Logging goes through `logging.getLogger("pgmpy")` like the reference (global_vars.py:7-29), so an application that
already filters or redirects pgmpy's log sees this engine's messages in the same place.
"""
from __future__ import annotations

import logging
import sys
from typing import Optional

logger = logging.getLogger("pgmpy")

_DTYPE_NAMES = {"float64": "float64", "double": "float64", "float32": "float32", "float": "float32", "single": "float32"}


def normalize_dtype(dtype) -> str:
    """'float64' / 'float32' from a name, a numpy dtype / scalar type or a torch dtype; ValueError otherwise."""
    if dtype is None:
        return "float64"
    name = getattr(dtype, "name", None) or getattr(dtype, "__name__", None) or str(dtype)
    name = str(name).replace("torch.", "").replace("numpy.", "").lower()
    if name not in _DTYPE_NAMES:
        raise ValueError(f"the B200 engine computes in float64 (default) or float32; got dtype {dtype!r}")
    return _DTYPE_NAMES[name]


class Config:
    """Same methods as pgmpy.global_vars.Config; BACKEND is always "b200" here."""

    def __init__(self):
        self.BACKEND = "b200"
        self.DTYPE = "float64"
        self.DEVICE: Optional[str] = None  # None = the current CUDA device
        self.SHOW_PROGRESS = True
        self._dtype_set = False
        # range-check evidence state indices that are handed over as CUDA tensors (one device reduction + host sync per
        # call); host arrays are always checked. Off by default: the kernels clamp into range for memory safety.
        self.validate_device_evidence = False

    def set_backend(self, backend: str = "b200", device: Optional[str] = None, dtype=None):
        if backend != "b200":
            raise ValueError(f"pgmpy_b200 has one backend, 'b200' (no numpy/torch/CPU execution path). Got: {backend}")
        self.set_device(device)
        self.set_dtype(dtype)

    def get_backend(self) -> str:
        return self.BACKEND

    def set_device(self, device: Optional[str] = None):
        if device is not None and not str(device).startswith("cuda"):
            raise ValueError(f"device must be 'cuda' or 'cuda:x' (there is no CPU execution path). Got: {device}")
        self.DEVICE = None if device is None else str(device)

    def get_device(self):
        return self.DEVICE

    def device_index(self) -> Optional[int]:
        if self.DEVICE is None or ":" not in self.DEVICE:
            return None
        return int(self.DEVICE.split(":", 1)[1])

    def set_dtype(self, dtype=None):
        self.DTYPE = normalize_dtype(dtype)
        self._dtype_set = dtype is not None

    def get_dtype(self) -> str:
        return self.DTYPE

    def set_show_progress(self, show_progress: bool):
        if not isinstance(show_progress, bool):
            raise ValueError(f"show_progress must be a boolean. Got: {show_progress}")
        self.SHOW_PROGRESS = show_progress

    def get_show_progress(self) -> bool:
        return self.SHOW_PROGRESS


config = Config()


def default_dtype() -> str:
    """dtype of a new inference object when the caller passes none (precedence in the module docstring)."""
    if config._dtype_set:
        return config.get_dtype()
    ref = sys.modules.get("pgmpy")
    ref_cfg = getattr(ref, "config", None) if ref is not None else None
    if ref_cfg is not None and hasattr(ref_cfg, "get_dtype"):
        try:
            return normalize_dtype(ref_cfg.get_dtype())
        except ValueError:
            logger.warning("pgmpy.config dtype %r is not supported by the B200 engine; using float64", ref_cfg.get_dtype())
    return config.get_dtype()
