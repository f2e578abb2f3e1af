"""ctypes binding of libpgx.so (C-ABI in include/pgx.h). Loading fails loudly: there is no fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGX_LIB", os.path.join(HERE, "libpgx.so"))  # PGX_LIB: tuning builds only

PGX_F64, PGX_F32 = 0, 1
MODE_AUTO, MODE_STEPWISE, MODE_FUSED = 0, 1, 2
OPT_MODE, OPT_FUSED_WARPS, OPT_USE_GRAPH, OPT_FUSED_KERNEL, OPT_STEP_KERNEL = 1, 2, 3, 4, 5
OPT_STAGE = 9
OPT_MMA = 10
OPT_TC32 = 11
INFO_N_STEPS, INFO_OUT_ELEMS, INFO_WS_ENTRIES, INFO_LAST_LAUNCHES, INFO_LAST_MODE, INFO_N_EV = 1, 2, 3, 4, 5, 6
INFO_LAST_VARIANT, INFO_N_LEVELS, INFO_LAST_GRAPH = 7, 8, 9
INFO_LAST_STAGED_STEPS = 10
INFO_IN_ELEMS = 11
INFO_LAST_TC_STEPS = 12
INFO_SPECIALIZED, INFO_SPEC_REGS, INFO_SPEC_SMEM, INFO_SPEC_COMPILE_MS, INFO_SPEC_LOADS, INFO_SPEC_FLOPS = 13, 14, 15, 16, 17, 18
FUSED_KERNELS = {"auto": 0, "generic": 1, "tables-smem": 2, "tables-global": 3, "specialized": 4}

EXPORTS = (
    "pgx_plan_create",
    "pgx_plan_destroy",
    "pgx_workspace_bytes",
    "pgx_run_batch",
    "pgx_run_batch_soft",
    "pgx_run_batch_multi",
    "pgx_plan_set_trace",
    "pgx_run_batch_mpe",
    "pgx_profile_steps",
    "pgx_profile_launches",
    "pgx_mm_pick",
    "pgx_plan_specialize",
    "pgx_spec_source",
    "pgx_plan_set_option",
    "pgx_plan_get_info",
    "pgx_evidence_reduce",
    "pgx_normalize",
    "pgx_argmax_rows",
    "pgx_batch_ld",
    "pgx_last_error",
    "pgx_abi_version",
)


class PlanDesc(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("dtype", C.c_int32),
        ("pool", C.POINTER(C.c_int32)),
        ("pool_words", C.c_int64),
        ("table_blob", C.c_void_p),
        ("table_entries", C.c_int64),
    ]


class PgxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"pgx error {code}: {msg}")
        self.code = code


_lib = None


def load():
    """Loads libpgx.so (building nothing: run `python -m pgmpy_b200.build` or __graft_entry__.build() first)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA engine is not built (python -m pgmpy_b200.build). "
            "pgmpy_b200 has no CPU execution path."
        )
    lib = C.CDLL(LIB_PATH)
    i32p = C.POINTER(C.c_int32)
    lib.pgx_plan_create.argtypes = [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]
    lib.pgx_plan_create.restype = C.c_int
    lib.pgx_plan_destroy.argtypes = [C.c_void_p]
    lib.pgx_plan_destroy.restype = None
    lib.pgx_workspace_bytes.argtypes = [C.c_void_p, C.c_int64]
    lib.pgx_workspace_bytes.restype = C.c_size_t
    lib.pgx_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]
    lib.pgx_run_batch.restype = C.c_int
    lib.pgx_run_batch_multi.argtypes = [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_int64), C.c_void_p]
    lib.pgx_run_batch_multi.restype = C.c_int
    lib.pgx_run_batch_soft.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]
    lib.pgx_run_batch_soft.restype = C.c_int
    lib.pgx_plan_set_trace.argtypes = [C.c_void_p, i32p, C.c_int64]
    lib.pgx_plan_set_trace.restype = C.c_int
    lib.pgx_run_batch_mpe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p]
    lib.pgx_run_batch_mpe.restype = C.c_int
    lib.pgx_profile_steps.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.POINTER(C.c_float), C.c_int32,
    ]
    lib.pgx_profile_steps.restype = C.c_int
    lib.pgx_profile_launches.argtypes = [
        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int64, C.c_void_p, C.POINTER(C.c_float), C.c_int32,
        i32p, C.c_int32, i32p,
    ]
    lib.pgx_profile_launches.restype = C.c_int
    lib.pgx_mm_pick.argtypes = [i32p, C.c_int32, C.c_int32, i32p, i32p, C.c_int64, C.POINTER(C.c_int64)]
    lib.pgx_mm_pick.restype = C.c_int
    lib.pgx_plan_specialize.argtypes = [C.c_void_p]
    lib.pgx_plan_specialize.restype = C.c_int
    lib.pgx_spec_source.argtypes = [C.POINTER(PlanDesc), C.c_int32, C.c_char_p, C.c_int64, C.POINTER(C.c_int64)]
    lib.pgx_spec_source.restype = C.c_int64
    lib.pgx_plan_set_option.argtypes = [C.c_void_p, C.c_int32, C.c_int64]
    lib.pgx_plan_set_option.restype = C.c_int
    lib.pgx_plan_get_info.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]
    lib.pgx_plan_get_info.restype = C.c_int
    lib.pgx_evidence_reduce.argtypes = [
        C.c_int32, C.c_void_p, C.c_int64, C.c_int32, i32p, i32p, C.c_int32, i32p, i32p, i32p, C.c_void_p, C.c_int32,
        C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
    ]
    lib.pgx_evidence_reduce.restype = C.c_int
    lib.pgx_normalize.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]
    lib.pgx_normalize.restype = C.c_int
    lib.pgx_argmax_rows.argtypes = [C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.pgx_argmax_rows.restype = C.c_int
    lib.pgx_batch_ld.argtypes = [C.c_int64]
    lib.pgx_batch_ld.restype = C.c_int64
    lib.pgx_last_error.argtypes = []
    lib.pgx_last_error.restype = C.c_char_p
    lib.pgx_abi_version.argtypes = []
    lib.pgx_abi_version.restype = C.c_int32
    if lib.pgx_abi_version() != 1:
        raise RuntimeError("libpgx.so ABI version mismatch")
    _lib = lib
    return lib


def check(code):
    if code != 0:
        raise PgxError(code, load().pgx_last_error().decode("utf-8", "replace"))
