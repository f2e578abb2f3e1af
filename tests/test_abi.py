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


def _mm_pick(rec, item_bytes=8, allow_mma=1):
    import ctypes as C

    lib = N.load()
    f = (C.c_int32 * 24)()
    n = C.c_int64()
    ptr = rec.ctypes.data_as(C.POINTER(C.c_int32))
    assert lib.pgx_mm_pick(ptr, item_bytes, allow_mma, f, None, 0, C.byref(n)) == 0
    if not f[0]:
        return None, None
    tabs = np.zeros(n.value, dtype=np.int32)
    assert lib.pgx_mm_pick(ptr, item_bytes, allow_mma, f, tabs.ctypes.data_as(C.POINTER(C.c_int32)), tabs.size, C.byref(n)) == 0
    keys = ("ok M N Z K lgTX lgTY TZ lgKC ntx nty ntz n_chunks n_stages stage_elems q_off p_const p_base q_base o_base "
            "n_active use_mma smem n_tiles").split()
    return dict(zip(keys, list(f))), tabs


@pytest.mark.parametrize("name", ["pathfinder", "diabetes", "munin"])
def test_mm_pick_tables_reproduce_the_step_definition(name):
    """Host-only (pgx_mm_pick, no GPU): for every step of a real junction-tree plan that k_contract_mm would take, the
    offset tables it indexes (flattened M/N/Z/K -> entries of P, Q and the output) must address exactly the entries the
    step record's mixed-radix definition does, the tiling must cover the index space and fit shared memory."""
    from pgmpy_b200.evidence import sample_evidence
    from pgmpy_b200.plan import OP_FIXED, STEP_FIXED
    from pgmpy_b200.planner import JTStructure, compile_jt_plan
    import pgmpy_b200 as px

    m = px.get_example_model(name)
    ev_vars, _ = sample_evidence(m, 1, 8, seed=1)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    pool = np.ascontiguousarray(plan.pool, dtype=np.int32)
    rng = np.random.default_rng(0)
    taken = mma = 0
    shapes = set()
    for si in range(plan.n_steps):
        rec = pool[int(pool[pool[10] + si]):]
        A, S, K2 = int(rec[0]), int(rec[1]), int(rec[2])
        f, tabs = _mm_pick(rec)
        if f is None:
            continue
        taken += 1
        mma += f["use_mma"]
        M, Nn, Z, K = f["M"], f["N"], f["Z"], f["K"]
        key = (M, Nn, Z, K, f["p_const"])
        if key in shapes and taken > 40:
            continue  # the same shape repeats per time slice: check a few of each
        shapes.add(key)
        odims = rec[STEP_FIXED:STEP_FIXED + A].astype(np.int64)
        sdims = rec[STEP_FIXED + A:STEP_FIXED + A + S].astype(np.int64)
        assert M * Nn * Z == int(np.prod(odims)) and K == int(np.prod(sdims))
        opw = OP_FIXED + A + S
        ops = [rec[STEP_FIXED + A + S + k * opw:][:opw].astype(np.int64) for k in range(2)]
        # the definition: every (o, s) pair -> (entry of op0, entry of op1), products summed per o
        o_idx = np.stack(np.unravel_index(np.arange(M * Nn * Z), odims), axis=1) if A else np.zeros((1, 0), np.int64)
        s_idx = np.stack(np.unravel_index(np.arange(K), sdims), axis=1)
        ent = []
        for op in ops:
            base = int(op[1])
            eo = o_idx @ op[OP_FIXED:OP_FIXED + A]
            es = s_idx @ op[OP_FIXED + A:OP_FIXED + A + S]
            ent.append(base + eo[:, None] + es[None, :])
        size = [int(e.max()) + 1 for e in ent]
        vals = [rng.random(n) for n in size]
        want = (vals[0][ent[0]] * vals[1][ent[1]]).sum(axis=1)
        # the kernel's addressing
        t = np.split(tabs, np.cumsum([M, M, Nn, Nn, K, K, Z, Z, Z, K]))
        xoffP, xoffO, yoffQ, yoffO, soffP, soffQ, zoffP, zoffQ, zoffO, soffPl, soffQl = t
        assert (soffQl == soffQ).all() and (soffPl == (0 if f["p_const"] else soffP)).all()  # ldb = 1 here
        swapped = f["p_base"] != int(ops[0][1]) or (int(ops[0][1]) == int(ops[1][1]) and False)
        pv, qv = (vals[1], vals[0]) if swapped else (vals[0], vals[1])
        pe = f["p_base"] + zoffP[:, None, None] + xoffP[None, :, None] + soffP[None, None, :]
        qe = f["q_base"] + zoffQ[:, None, None] + yoffQ[None, :, None] + soffQ[None, None, :]
        got = np.einsum("zxk,zyk->zxy", pv[pe], qv[qe])
        oe = zoffO[:, None, None] + xoffO[None, :, None] + yoffO[None, None, :]
        assert sorted(oe.reshape(-1)) == list(range(M * Nn * Z))  # a bijection onto the output entries
        flat = np.empty(M * Nn * Z)
        flat[oe.reshape(-1)] = got.reshape(-1)
        np.testing.assert_allclose(flat, want, rtol=1e-12)
        # tiling: covers the index space, fits the budget, one register block per active warp
        TX, TY = 1 << f["lgTX"], 1 << f["lgTY"]
        assert f["ntx"] * TX >= M and f["nty"] * TY >= Nn and f["ntz"] * f["TZ"] >= Z
        assert f["n_active"] == f["TZ"] * TX * TY // 32 <= 16 and 1 <= f["n_stages"] <= 8
        assert f["smem"] <= 227 * 1024 and (f["n_chunks"] << f["lgKC"]) >= K
    assert taken > 0
    if name == "diabetes":
        assert mma > 0  # CPT-times-message steps go to the tensor-core consumer
