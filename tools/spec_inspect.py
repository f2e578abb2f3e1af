"""Generate (and optionally NVRTC-compile) the plan-specialised kernel of a model on the CPU: source, statistics, SASS.

    python tools/spec_inspect.py alarm [k] [--compile] [--lib path/to/libpgx.so] [--out file.cu] [--dtype float32]
No GPU needed (NVRTC cross-compiles)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import pgmpy_b200 as px
from pgmpy_b200 import _native as N
from pgmpy_b200.evidence import sample_evidence
from pgmpy_b200.planner import JTStructure, compile_jt_plan


from pgmpy_b200.specialize import spec_source as _spec_source


def spec_source(lib, plan, dtype="float64", compile=0):
    return _spec_source(plan, dtype, compile, lib)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    name = args[0] if args else "alarm"
    k = int(args[1]) if len(args) > 1 else 5
    libpath = sys.argv[sys.argv.index("--lib") + 1] if "--lib" in sys.argv else N.LIB_PATH
    dtype = sys.argv[sys.argv.index("--dtype") + 1] if "--dtype" in sys.argv else "float64"
    lib = C.CDLL(libpath)
    m = px.get_example_model(name)
    jt = JTStructure.from_model(m)
    ev_vars, _ = sample_evidence(m, 1, k, seed=1)
    plan = compile_jt_plan(jt, ev_vars)
    src, st = spec_source(lib, plan, dtype, 0)
    print(name, "ws_entries (plan)", plan.ws_entries, st, "source bytes", len(src))
    if "--out" in sys.argv:
        open(sys.argv[sys.argv.index("--out") + 1], "wb").write(src)
    if "--compile" in sys.argv:
        cubin, st = spec_source(lib, plan, dtype, 2)
        print("compiled:", st)
        open("/tmp/spec_kernel.cubin", "wb").write(cubin)
        os.system("cuobjdump -res-usage /tmp/spec_kernel.cubin | grep -i -A1 k_plan_spec | tail -2")
        os.system("cuobjdump -sass /tmp/spec_kernel.cubin | grep -E '^\\s+/\\*[0-9a-f]{4,}\\*/' | awk '{print $2}' | sed 's/\\..*//' | sort | uniq -c | sort -rn | head -25")
