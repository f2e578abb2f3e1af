"""Plan-specialised kernel (pgmpy_b200/csrc/pgx_spec.cu), CPU side: the generator's output for real junction-tree
plans is (a) compiled for sm_100a by NVRTC without a GPU and (b) executed on the CPU — the very same source, one lane
at a time through tests/hostsim/spec_host.cpp — and compared with the numpy plan interpreter (oracle/plan_exec.py,
test infrastructure) on seeded evidence. Covers indexing, lifetime packing of the work tables, constant folding of
evidence-independent messages, ragged batches, fp32 mode."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import pgmpy_b200 as px
from pgmpy_b200 import _native as N
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan
from oracle.plan_exec import run_plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from spec_inspect import spec_source  # noqa: E402


def _host_run(src: bytes, tmp_path, blob, states, out_elems, dtype):
    gen = tmp_path / "gen.cu"
    gen.write_bytes(src)
    so = tmp_path / "spec_host.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", f'-DPGX_GENERATED="{gen}"', "-o", str(so),
                    os.path.join(ROOT, "tests", "hostsim", "spec_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    np_dt = np.float64 if dtype == "float64" else np.float32
    blob = np.ascontiguousarray(blob, dtype=np_dt)
    ev = np.ascontiguousarray(states, dtype=np.int32)
    B = ev.shape[0]
    out = np.full((B, out_elems), -7.0, dtype=np_dt)
    lib.spec_host_run(C.c_void_p(blob.ctypes.data), C.c_void_p(ev.ctypes.data), C.c_void_p(out.ctypes.data), C.c_longlong(B))
    return out


@pytest.mark.parametrize("name,k,dtype,distribute", [
    ("asia", 2, "float64", "ss"), ("child", 4, "float64", "ss"), ("alarm", 5, "float64", "ss"), ("alarm", 0, "float64", "ss"),
    ("alarm", 5, "float32", "ss"), ("sachs", 3, "float64", "ss"),
    # belief-update plans: divide steps (sigma / mu with 0 / 0 -> 0, ExactInference.py:788-805)
    ("alarm", 5, "float64", "divide"), ("alarm", 5, "float64", "belief"), ("hepar2", 5, "float64", "auto"),
    ("win95pts", 5, "float64", "auto")])
def test_generated_source_on_the_cpu_matches_the_plan_interpreter(name, k, dtype, distribute, tmp_path):
    lib = N.load()
    m = px.get_example_model(name)
    B = 70  # two full rows of 32 and a ragged one
    ev_vars, states = sample_evidence(m, B, k, seed=11)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars, distribute=distribute)
    src, st = spec_source(lib, plan, dtype, 0)
    assert st["smem_bytes"] <= 226 * 1024
    assert st["loads"] < st["terms"] * 3  # distinct elements, not one load per factor of every term
    states = states.reshape(B, -1) if k else np.zeros((B, 0), np.int32)
    got = _host_run(src, tmp_path, plan.const_blob, states, plan.out_elems, dtype)
    want = run_plan(plan.pool, plan.const_blob, states if k else np.zeros((B, 0), np.int32))
    err = np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))
    assert err <= (1e-12 if dtype == "float64" else 1e-5), err


def test_generated_source_compiles_for_sm_100a_without_a_gpu():
    lib = N.load()
    m = px.get_example_model("alarm")
    ev_vars, _ = sample_evidence(m, 1, 5, seed=1)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars)
    cubin, st = spec_source(lib, plan, "float64", 2)
    assert cubin[:4] == b"\x7fELF" and st["cubin_bytes"] == len(cubin) and st["compile_ms"] > 0


def test_compiled_kernels_are_cached_on_disk_when_asked(tmp_path, monkeypatch):
    """PGX_SPEC_CACHE_DIR: the cubin of a generated source is stored under its hash; the second compile is a file read."""
    lib = N.load()
    m = px.get_example_model("child")
    ev_vars, _ = sample_evidence(m, 1, 4, seed=1)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars, distribute="ss")
    monkeypatch.setenv("PGX_SPEC_CACHE_DIR", str(tmp_path))
    cubin1, st1 = spec_source(lib, plan, "float64", 2)
    files = list(tmp_path.glob("pgx_spec_sm100a_*.cubin"))
    assert len(files) == 1 and files[0].read_bytes() == cubin1
    cubin2, st2 = spec_source(lib, plan, "float64", 2)
    assert cubin2 == cubin1 and st2["compile_ms"] <= max(20, st1["compile_ms"] // 5)
    # a different evidence signature is a different source: its own entry
    ev2, _ = sample_evidence(m, 1, 3, seed=2)
    spec_source(lib, compile_jt_plan(JTStructure.from_model(m), ev2, distribute="ss"), "float64", 2)
    assert len(list(tmp_path.glob("pgx_spec_sm100a_*.cubin"))) == 2


def test_plans_the_generator_refuses():
    lib = N.load()
    m = px.get_example_model("alarm")
    ev_vars, _ = sample_evidence(m, 1, 5, seed=1)
    plan = compile_jt_plan(JTStructure.from_model(m), ev_vars, reduce_max=True)
    with pytest.raises(RuntimeError, match="max-reduce"):
        spec_source(lib, plan, "float64", 0)
    big = px.get_example_model("pathfinder")
    ev_vars, _ = sample_evidence(big, 1, 8, seed=1)
    with pytest.raises(RuntimeError, match="too large|shared memory"):
        spec_source(lib, compile_jt_plan(JTStructure.from_model(big), ev_vars), "float64", 0)
