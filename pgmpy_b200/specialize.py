"""Host-side access to the plan specialiser (csrc/pgx_spec.cu) without a GPU: the generated CUDA source of a plan, its
statistics (product terms, loads, fp instructions, shared memory) and, optionally, the NVRTC-compiled cubin. Used by the
planner to choose the plan variant with the fewest multiply-adds when a plan is going to be specialised, by
tools/spec_inspect.py and by the CPU tests."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N

STAT_KEYS = ("terms", "terms_kept", "loads", "flops", "ws_entries", "smem_bytes", "compile_ms", "cubin_bytes")


def spec_source(plan, dtype: str = "float64", compile: int = 0, lib=None):
    """(bytes, stats): the generated source (compile = 0 / 1) or the cubin (compile = 2) of `plan`'s specialised kernel.
    Raises RuntimeError with the generator's reason when the plan cannot be specialised."""
    lib = lib or N.load()
    pool = np.ascontiguousarray(plan.pool, dtype=np.int32)
    blob = np.ascontiguousarray(plan.const_blob, dtype=np.float64 if dtype == "float64" else np.float32)
    if blob.size == 0:
        blob = np.zeros(1, dtype=blob.dtype)
    desc = N.PlanDesc(1, N.PGX_F64 if dtype == "float64" else N.PGX_F32, pool.ctypes.data_as(C.POINTER(C.c_int32)), pool.size,
                      C.c_void_p(blob.ctypes.data), blob.size)
    stats = (C.c_int64 * 8)()
    lib.pgx_spec_source.restype = C.c_int64
    lib.pgx_spec_source.argtypes = [C.POINTER(N.PlanDesc), C.c_int32, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    need = lib.pgx_spec_source(C.byref(desc), 0 if compile != 2 else 2, None, 0, stats) if compile == 0 else 0
    cap = max(int(need), 0) + 16 if compile == 0 else 64 << 20
    buf = C.create_string_buffer(max(cap, 4096))
    n = lib.pgx_spec_source(C.byref(desc), compile, buf, len(buf), stats)
    if n < 0:
        raise RuntimeError(f"pgx_spec_source: {n}: {buf.value.decode(errors='replace')[:2000]}")
    return buf.raw[:n], dict(zip(STAT_KEYS, list(stats)))


def spec_flops(plan, dtype: str = "float64"):
    """fp instructions per row of 32 evidence sets the generator emits for `plan`, or None when it refuses the plan."""
    try:
        return spec_source(plan, dtype, 0)[1]["flops"]
    except Exception:
        return None
