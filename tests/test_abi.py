"""The C-ABI library loads on a CPU-only box and exports every symbol include/pgx.h declares; descriptor
validation (which runs before any CUDA call) rejects malformed plans. No compute calls here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import pgmpy_b200 as px
from pgmpy_b200 import _native as N
from pgmpy_b200.planner import JTStructure, compile_jt_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported():
    header = open(os.path.join(ROOT, "include", "pgx.h")).read()
    declared = set(re.findall(r"\b(pgx_[a-z_0-9]+)\s*\(", header))
    assert declared == set(N.EXPORTS)
    lib = N.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pgx_abi_version() == 1


def test_batch_leading_dimension():
    lib = N.load()
    assert [lib.pgx_batch_ld(b) for b in (1, 2, 3, 17, 32, 33, 1000)] == [1, 2, 4, 32, 32, 64, 1024]


def _create(pool, entries=1 << 20, dtype=0, blob=0x1000):
    lib = N.load()
    pool = np.ascontiguousarray(pool, dtype=np.int32)
    desc = N.PlanDesc(1, dtype, pool.ctypes.data_as(C.POINTER(C.c_int32)), pool.size, C.c_void_p(blob), entries)
    h = C.c_void_p()
    code = lib.pgx_plan_create(C.byref(desc), C.byref(h))
    return code, lib.pgx_last_error().decode(), h


def test_plan_validation_rejects_malformed_pools():
    m = px.get_example_model("asia")
    plan = compile_jt_plan(JTStructure.from_model(m), ["xray"])
    good = plan.pool.copy()
    bad = good.copy()
    bad[0] = 7
    code, msg, _ = _create(bad)
    assert code == -1 and "magic" in msg
    code, msg, _ = _create(good[:10])
    assert code == -1
    code, msg, _ = _create(good, entries=3)  # table blob too small
    assert code == -2
    bad = good.copy()
    first = bad[bad[10]]  # first step record
    bad[first + 8] = int(plan.ws_entries)  # output offset beyond the workspace
    code, msg, _ = _create(bad)
    assert code == -2 and "workspace" in msg
    bad = good.copy()
    bad[first + 2] = 99  # operand count
    assert _create(bad)[0] == -5
    code, msg, _ = _create(good, dtype=7)
    assert code == -1


def test_valid_plan_needs_a_gpu_and_never_falls_back():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    m = px.get_example_model("asia")
    plan = compile_jt_plan(JTStructure.from_model(m), ["xray"])
    code, msg, _ = _create(plan.pool, entries=plan.const_blob.size)
    assert code == -4, (code, msg)  # PGX_ERR_CUDA: no device
    from pgmpy_b200.inference import VariableElimination

    with pytest.raises(RuntimeError, match="no CPU execution path"):
        VariableElimination(m).query(["lung"], evidence={"xray": "yes"})


def test_run_batch_argument_checks():
    lib = N.load()
    assert lib.pgx_run_batch(None, None, None, None, 0, 1, None) == -1
    assert lib.pgx_workspace_bytes(None, 10) == 0


def test_stage_pick_finds_the_gemm_shape_of_a_two_operand_step():
    """Host-only tile picker of the TMA-staged GEMM-tile kernel (pgx_stage_pick, no GPU needed): for
    out[x, y] = sum_s P[x, s] * Q[s, y] it must choose form (X-only, Y-only) with the two output axes as tile axes."""
    import ctypes as C

    from pgmpy_b200 import _native as N
    from pgmpy_b200.plan import PlanBuilder

    card = {"x": 21, "y": 55, "s": 96}
    b = PlanBuilder(card, [])
    p0 = b.add_const(["x", "s"], np.ones((21, 96)))
    q0 = b.add_const(["s", "y"], np.ones((96, 55)))
    p = b.contract([p0], ["x", "s"])  # work tables (batch dependent in a real plan)
    q = b.contract([q0], ["s", "y"])
    out = b.contract([p, q], ["x", "y"])
    b.emit(out, False)
    plan = b.finalize()
    pool = np.ascontiguousarray(plan.pool, dtype=np.int32)
    rec_off = int(pool[pool[10] + plan.n_steps - 1])  # last step = the contraction
    lib = N.load()
    f = (C.c_int32 * 12)()
    sm = C.c_int64()
    rec = pool[rec_off:]
    assert lib.pgx_stage_pick(rec.ctypes.data_as(C.POINTER(C.c_int32)), 8, f, C.byref(sm)) == 0
    ok, ax, ay, bx, by, ntx, nty, tiles, sc, swap, form, stage_elems = list(f)
    assert ok == 1 and {ax, ay} == {0, 1} and form == 1 * 4 + 2
    assert 1 <= bx * by <= 8 and ntx * 4 * bx >= card["x" if ax == 0 else "y"] and 1 <= sc <= 96
    assert 0 < sm.value <= 200 * 1024 and tiles == ntx * nty
