"""numpy interpreter for packed contraction plans (pgmpy_b200.plan word-pool layout).

TEST INFRASTRUCTURE. It parses the same int32 pool the CUDA engine receives and evaluates every step with
plain numpy gathers, so planner bugs show up on the CPU before a kernel runs, and kernel bugs can be
bisected step by step. The product package never imports this file.
"""
import numpy as np

from pgmpy_b200 import plan as P


def _i64(pool, at):
    return P.lohi_to_int(pool[at], pool[at + 1])


def parse(pool):
    pool = np.asarray(pool, dtype=np.int64)
    assert pool[0] == P.MAGIC and pool[1] == P.VERSION
    hdr = dict(
        n_ev=int(pool[2]),
        n_steps=int(pool[3]),
        n_segs=int(pool[4]),
        out_elems=int(pool[5]),
        ws_entries=_i64(pool, 6),
        const_entries=_i64(pool, 8),
        step_index_off=int(pool[10]),
        segs_off=int(pool[11]),
    )
    steps = []
    for s in range(hdr["n_steps"]):
        base = int(pool[hdr["step_index_off"] + s])
        A, S, K, flags = (int(x) for x in pool[base : base + 4])
        st = dict(
            A=A, S=S, K=K, flags=flags,
            out_size=_i64(pool, base + 4), sum_size=_i64(pool, base + 6), out_off=_i64(pool, base + 8),
            level=int(pool[base + 10]),
            out_dims=[int(x) for x in pool[base + 12 : base + 12 + A]],
            sum_dims=[int(x) for x in pool[base + 12 + A : base + 12 + A + S]],
            ops=[],
        )
        opw = P.OP_FIXED + A + S
        for k in range(K):
            ob = base + P.STEP_FIXED + A + S + k * opw
            kind = int(pool[ob])
            n_ev = int(pool[ob + 3])
            evo = base + int(pool[ob + 4])
            st["ops"].append(
                dict(
                    kind=kind & 0xFF, divisor=bool(kind & 0x100), offset=_i64(pool, ob + 1),
                    so=[int(x) for x in pool[ob + 6 : ob + 6 + A]],
                    ss=[int(x) for x in pool[ob + 6 + A : ob + 6 + A + S]],
                    ev=[(int(pool[evo + 2 * j]), int(pool[evo + 2 * j + 1])) for j in range(n_ev)],
                )
            )
        steps.append(st)
    segs = []
    for g in range(hdr["n_segs"]):
        sb = hdr["segs_off"] + g * P.SEG_WORDS
        segs.append(dict(off=_i64(pool, sb), size=int(pool[sb + 2]), out_off=int(pool[sb + 3]), flags=int(pool[sb + 4])))
    inputs = []
    ib = int(pool[15])
    hdr["in_elems"] = 0
    if ib:
        hdr["in_elems"] = int(pool[ib + 1])
        for j in range(int(pool[ib])):
            at = ib + 2 + 4 * j
            inputs.append(dict(off=_i64(pool, at), size=int(pool[at + 2]), in_off=int(pool[at + 3])))
    hdr["inputs"] = inputs
    return hdr, steps, segs


def _grid(dims, strides):
    """flat offsets for a mixed-radix index space, row-major (last axis fastest)."""
    off = np.zeros((1,), dtype=np.int64)
    for d, s in zip(dims, strides):
        off = (off[:, None] + (np.arange(d, dtype=np.int64) * s)[None, :]).reshape(-1)
    return off


def run_plan(pool, const_blob, ev_states, dtype=np.float64, return_workspace=False, soft=None):
    """Returns out[B, out_elems]. ev_states: int [B, n_ev]; soft: [B, in_elems] rows for the plan's input tables."""
    hdr, steps, segs = parse(pool)
    ev_states = np.asarray(ev_states, dtype=np.int64).reshape(-1, max(hdr["n_ev"], 0)) if hdr["n_ev"] else np.zeros(
        (np.asarray(ev_states).shape[0], 0), dtype=np.int64
    )
    B = ev_states.shape[0]
    const = np.asarray(const_blob, dtype=dtype)
    ws = np.zeros((hdr["ws_entries"], B), dtype=dtype)
    bidx = np.arange(B)
    if hdr["inputs"]:
        soft = np.asarray(soft, dtype=dtype).reshape(B, hdr["in_elems"])
        for inp in hdr["inputs"]:
            ws[inp["off"] : inp["off"] + inp["size"], :] = soft[:, inp["in_off"] : inp["in_off"] + inp["size"]].T
    for st in steps:
        O, S_ = st["out_size"], st["sum_size"]
        num = np.ones((B, O, S_), dtype=dtype)
        den = None
        for op in st["ops"]:
            idx = op["offset"] + _grid(st["out_dims"], op["so"])[:, None] + _grid(st["sum_dims"], op["ss"])[None, :]
            evoff = np.zeros(B, dtype=np.int64)
            for slot, stride in op["ev"]:
                evoff += ev_states[:, slot] * stride
            full = idx[None, :, :] + evoff[:, None, None]
            if op["kind"] == P.KIND_CONST:
                vals = const[full]
            else:
                vals = ws[full, bidx[:, None, None]]
            if op["divisor"]:
                d = vals[:, :, 0]
                den = d if den is None else den * d
            else:
                num = num * vals
        red = num.max(axis=2) if st["flags"] & P.FLAG_MAX else num.sum(axis=2)
        if den is not None:
            with np.errstate(divide="ignore", invalid="ignore"):
                red = red / den
            red = np.where(np.isnan(red), 0.0, red)
        ws[st["out_off"] : st["out_off"] + O, :] = red.T
    out = np.zeros((B, hdr["out_elems"]), dtype=dtype)
    for sg in segs:
        seg = ws[sg["off"] : sg["off"] + sg["size"], :].T
        if sg["flags"] & P.SEG_NORMALIZE:
            with np.errstate(divide="ignore", invalid="ignore"):
                seg = seg / seg.sum(axis=1, keepdims=True)
        out[:, sg["out_off"] : sg["out_off"] + sg["size"]] = seg
    if return_workspace:
        return out, ws
    return out


def mpe_traceback(trace, ws):
    """numpy restatement of k_mpe_traceback: walks the descriptor compile_jt_mpe_plan returns over the workspace of
    run_plan(..., return_workspace=True) and returns the argmax assignment, int [B, n_columns]."""
    trace = np.asarray(trace, dtype=np.int64)
    n_cliques, n_cols = int(trace[0]), int(trace[1])
    B = ws.shape[1]
    assign = np.zeros((B, n_cols), dtype=np.int64)
    at = 2
    for _ in range(n_cliques):
        off = P.lohi_to_int(trace[at], trace[at + 1])
        n_ax = int(trace[at + 2])
        axes = trace[at + 3 : at + 3 + 4 * n_ax].reshape(n_ax, 4)
        at += 3 + 4 * n_ax
        new = [a for a in axes if a[3]]
        dims = [int(a[1]) for a in new]
        rel = _grid(dims, [int(a[2]) for a in new])  # C-order over the axes assigned here
        for b in range(B):
            base = off + sum(int(assign[b, a[0]]) * int(a[2]) for a in axes if not a[3])
            vals = ws[base + rel, b]
            k = int(np.argmax(vals))  # first maximum
            for a, d in zip(new, np.unravel_index(k, dims) if dims else ()):
                assign[b, a[0]] = d
    return assign
