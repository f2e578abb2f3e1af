"""TEST INFRASTRUCTURE — not product code.

`oracle/` holds the CPU checker for the B200 exact-inference path:

* `ref_loader.py`   imports the real, unmodified reference (pgmpy 1.0.0 at /root/reference) through
                    the stand-in packages in `shims/` (only possible in the build container);
* `pgm_oracle.py`   a numpy restatement of the reference's algorithm (factor algebra, pruning,
                    variable elimination, belief propagation), each function citing the reference
                    file:line it follows; pinned against the reference's own known-answer tests and
                    against outputs of the reference itself (tests/golden/, made by `make_golden.py`);
* `plan_exec.py`    a numpy interpreter for the engine's contraction plans (validates the planner
                    on CPU before a kernel ever runs).

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import anything from here. The product package `pgmpy_b200` never does.
"""
