// pgx_spec.cu — plan-specialised whole-plan kernel: code generator, NVRTC compile, driver-API load and launch.
// See pgx_spec.h for the idea and DESIGN.md 4.2 for the measurements behind every choice below. The generated kernel
// computes what k_plan_fused2 computes (same step list, same summation order inside a step) the way a compiler can
// when every address and every batch-invariant value is known:
//
//   * one warp = one row of 32 evidence sets (lane = evidence set) runs the WHOLE plan; lane l only ever reads what
//     lane l wrote, so a work-table entry is a local variable (register, or a local-memory spill where ptxas decides):
//     no offset tables, no index arithmetic, no barrier, no work tables in memory at all;
//   * batch-invariant operands (CPTs / clique potentials that hold no observed variable) are immediates; the
//     coefficients of a term are folded on the host and terms with a zero coefficient are not emitted; entries without
//     any evidence-dependent factor (messages out of subtrees that hold no evidence) are evaluated here and folded
//     into their consumers; a CPT entry indexed by one observed variable is a select among immediates;
//   * only the output entries live in shared memory, laid out as the posterior rows themselves: the kernel is
//     persistent and every row leaves as ONE TMA bulk store (cp.async.bulk shared -> global) that drains while the
//     next row computes; the per-variable normalisation takes one branch per row;
//   * the variants that were measured and lost (several warps per row with level barriers, two warps along a cut of
//     the junction tree, lock-stepped rows, partial sums, deferred marginals) stay behind PGX_SPEC_* environment knobs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <vector>

#include "pgx_spec.h"
#include "pgx_step.cuh"

namespace pgx {

namespace {

// Shared-memory layout of a row's work tables: [entry][32] with lane l's element of entry e at position l ^ (e & 15).
// Every access of the compute phase is a whole aligned 32-element row (2 wavefronts for doubles whatever the
// permutation — a pitch of 33 would straddle three 128-byte lines and cost 3), and the transposed reads of the output
// stage (lanes = consecutive entries, same evidence set) fall into different banks. e is a compile-time constant, so
// the permutation costs 16 lane-base registers (lane ^ 0 .. lane ^ 15) and no instruction.
constexpr int P = 32;
constexpr int SWZ = 15;

struct Op {
    bool work = false;
    int64_t base = 0;
    std::vector<std::pair<int, int>> pairs;  // (evidence slot, stride)
    std::vector<int> ostr, sstr;
    int owner = -1;  // step whose output this work operand reads
    bool div = false;
};
struct Step {
    int A = 0, S = 0, K = 0, flags = 0, level = 0;
    int64_t out_size = 0, sum_size = 0, out_off = 0;
    std::vector<int> odims, sdims;
    std::vector<Op> ops;
};

struct Emitter {
    std::string s;
    void line(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        const int n = vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        if (n >= (int)sizeof buf) {
            std::vector<char> big((size_t)n + 1);
            va_start(ap, fmt);
            vsnprintf(big.data(), big.size(), fmt, ap);
            va_end(ap);
            s += big.data();
        } else {
            s += buf;
        }
        s += '\n';
    }
};

std::string lit(double v, bool f32) {
    char buf[64];
    if (v == 0.0) return f32 ? "0.0f" : "0.0";
    if (std::isinf(v)) return v > 0 ? (f32 ? "__int_as_float(0x7f800000)" : "__longlong_as_double(0x7ff0000000000000LL)")
                                    : (f32 ? "__int_as_float(0xff800000)" : "__longlong_as_double(0xfff0000000000000LL)");
    snprintf(buf, sizeof buf, f32 ? "%af" : "%a", v);
    return buf;
}

}  // namespace

bool pgx_spec_generate(const int32_t* pool, int64_t pool_words, const void* host_blob, int64_t table_entries, int dtype,
                       std::string& source, SpecStats& stats, std::string& why, int warps) {
    (void)pool_words;
    const int G = warps < 1 ? 1 : (warps > 16 ? 16 : warps);
    const bool f32 = dtype == 1;
    const size_t elem = f32 ? 4 : 8;
    const int n_ev = pool[2], n_steps = pool[3], n_segs = pool[4], out_elems = pool[5];
    const int64_t ws_entries = ld_i64(pool + 6);
    const int32_t* index = pool + pool[10];
    const int32_t* segs = pool + pool[11];
    const int32_t* ev_card = pool + pool[14];
    if (pool[15] != 0) {
        why = "plan has batch-dependent input tables";
        return false;
    }
    if (n_steps < 1 || out_elems < 1) {
        why = "empty plan";
        return false;
    }
    auto cval = [&](int64_t i) -> double {
        if (i < 0 || i >= table_entries) return 0.0;
        return f32 ? (double)((const float*)host_blob)[i] : ((const double*)host_blob)[i];
    };
    // ---- parse
    std::vector<Step> steps(n_steps);
    int64_t terms = 0;
    for (int si = 0; si < n_steps; ++si) {
        const int32_t* r = pool + index[si];
        Step& st = steps[si];
        st.A = r[0], st.S = r[1], st.K = r[2], st.flags = r[3], st.level = r[10];
        st.out_size = ld_i64(r + 4), st.sum_size = ld_i64(r + 6), st.out_off = ld_i64(r + 8);
        if (st.flags & ~FLAG_DIV) {
            why = "plan has max-reduce steps";
            return false;
        }
        const int opw = OP_FIXED + st.A + st.S;
        const int32_t* odims = r + STEP_FIXED;
        const int32_t* sdims = odims + st.A;
        const int32_t* ops = sdims + st.S;
        st.odims.assign(odims, odims + st.A);
        st.sdims.assign(sdims, sdims + st.S);
        st.ops.resize(st.K);
        for (int k = 0; k < st.K; ++k) {
            const int32_t* op = ops + k * opw;
            Op& o = st.ops[k];
            o.div = (op[0] & 0x100) != 0;  // trailing operands of a divide step: out = (sum of products) / prod(divisors)
            o.work = (op[0] & 0xFF) == 1;
            o.base = ld_i64(op + 1);
            for (int j = 0; j < op[3]; ++j) o.pairs.push_back({r[op[4] + 2 * j], r[op[4] + 2 * j + 1]});
            o.ostr.assign(op + OP_FIXED, op + OP_FIXED + st.A);
            o.sstr.assign(op + OP_FIXED + st.A, op + OP_FIXED + st.A + st.S);
        }
        terms += st.out_size * st.sum_size;
    }
    if (terms > 120000) {
        why = "plan too large for straight-line code (" + std::to_string(terms) + " product terms)";
        return false;
    }
    // ---- lifetime packing of the work tables. Objects = step outputs (the planner may already have reused offsets, so
    // "the object at entry e" depends on the step that asks: owners are tracked in program order); an object dies
    // after its last reader and its slots are handed to later outputs (first fit).
    struct Obj {
        int64_t off, size, new_off;
        int last;
    };
    std::vector<Obj> objs;
    std::vector<int> seg_owner(n_segs, -1);
    bool pack = true;
    std::string nopack;
    {
        std::vector<int> cur(ws_entries, -1);
        for (int si = 0; si < n_steps; ++si) objs.push_back({steps[si].out_off, steps[si].out_size, steps[si].out_off, si});
        auto owner = [&](int64_t lo, int64_t hi) -> int {
            if (lo < 0 || hi >= ws_entries) return -1;
            const int ow = cur[lo];
            for (int64_t e = lo; e <= hi; ++e)
                if (cur[e] != ow) return -1;
            return ow;
        };
        for (int si = 0; si < n_steps && pack; ++si) {
            Step& st = steps[si];
            for (Op& o : st.ops) {
                if (!o.work) continue;
                int64_t lo = o.base, hi = o.base;
                for (int a = 0; a < st.A; ++a) (o.ostr[a] < 0 ? lo : hi) += (int64_t)o.ostr[a] * (st.odims[a] - 1);
                for (int a = 0; a < st.S; ++a) (o.sstr[a] < 0 ? lo : hi) += (int64_t)o.sstr[a] * (st.sdims[a] - 1);
                for (auto& pr : o.pairs) (pr.second < 0 ? lo : hi) += (int64_t)pr.second * (ev_card[pr.first] - 1);
                const int ow = owner(lo, hi);
                if (ow < 0) {
                    pack = false;
                    nopack = "step " + std::to_string(si) + " reads [" + std::to_string(lo) + "," + std::to_string(hi) + "]: not one step output";
                    break;
                }
                o.owner = ow;
                objs[ow].last = std::max(objs[ow].last, si);
            }
            if (st.out_off < 0 || st.out_off + st.out_size > ws_entries) {
                why = "step output outside the workspace";
                return false;
            }
            for (int64_t e = st.out_off; e < st.out_off + st.out_size; ++e) cur[e] = si;
        }
        for (int g = 0; g < n_segs && pack; ++g) {
            const int32_t* sg = segs + g * SEG_WORDS;
            const int64_t lo = ld_i64(sg), hi = lo + sg[2] - 1;
            const int ow = owner(lo, hi);
            if (ow < 0) {
                pack = false, nopack = "segment " + std::to_string(g) + " is not inside one step output";
            } else {
                seg_owner[g] = ow;
                objs[ow].last = n_steps;  // lives to the output stage
            }
        }
    }
    int64_t packed_entries = ws_entries;
    if (pack) {
        std::vector<std::pair<int64_t, int64_t>> free_list;  // (off, size), sorted by off
        int64_t top = 0;
        std::vector<std::vector<int>> dying(n_steps + 1);
        for (int i = 0; i < n_steps; ++i) dying[std::min(objs[i].last, n_steps)].push_back(i);
        for (int si = 0; si < n_steps; ++si) {
            Obj& o = objs[si];
            bool placed = false;
            for (size_t f = 0; f < free_list.size(); ++f) {
                if (free_list[f].second >= o.size) {
                    o.new_off = free_list[f].first;
                    free_list[f].first += o.size;
                    free_list[f].second -= o.size;
                    if (free_list[f].second == 0) free_list.erase(free_list.begin() + f);
                    placed = true;
                    break;
                }
            }
            if (!placed) {
                o.new_off = top;
                top += o.size;
            }
            // the slots of objects whose last reader is step si are free from the NEXT step on (a step never writes
            // into one of its own operands)
            for (int d : dying[si]) free_list.push_back({objs[d].new_off, objs[d].size});
            std::sort(free_list.begin(), free_list.end());
            for (size_t f = 1; f < free_list.size();) {  // coalesce
                if (free_list[f - 1].first + free_list[f - 1].second == free_list[f].first) {
                    free_list[f - 1].second += free_list[f].second;
                    free_list.erase(free_list.begin() + f);
                } else {
                    ++f;
                }
            }
        }
        packed_entries = std::max<int64_t>(top, 1);
        // several warps per row run the steps of a level concurrently: slots may only be reused across levels, which is
        // what the planner's own allocation guarantees (the stepwise path has the same need) -> keep it then
        if (G > 1 || packed_entries >= ws_entries) {
            for (Obj& o : objs) o.new_off = o.off;
            packed_entries = ws_entries;
        }
    }
    // One warp per row (G == 1): a work entry is read only by the thread that wrote it, so it does not need memory at
    // all — it becomes a local variable (register, or a local-memory spill where the compiler decides) unless the
    // output stage reads it transposed (segments) or a step indexes it with evidence (dynamic address). Shared memory
    // then holds the output entries only (alarm: 91 of 267), which is what lets 8 rows be resident per SM instead of 3.
    std::vector<char> in_smem(n_steps, 1);
    if (G == 1 && pack) {
        std::fill(in_smem.begin(), in_smem.end(), 0);
        for (int g = 0; g < n_segs; ++g) in_smem[seg_owner[g]] = 1;
        for (const Step& st : steps)
            for (const Op& o : st.ops)
                if (o.work && !o.pairs.empty()) in_smem[o.owner] = 1;
        int64_t top = 0;
        for (int si = 0; si < n_steps; ++si)
            if (in_smem[si]) objs[si].new_off = top, top += objs[si].size;
        packed_entries = std::max<int64_t>(top, 1);
    }
    // Staged-output mode (the default for one warp per row): the output entries are kept in shared memory in the layout
    // of the posterior rows themselves, [evidence set][out_elems] (pitch made odd: conflict free for lane = evidence
    // set), so a row's 32 x out_elems block of out[] is ONE contiguous copy: a TMA bulk store (cp.async.bulk
    // shared -> global) issued by one lane. The kernel is persistent (a CTA walks rows blockIdx.x, + gridDim.x, ...):
    // the store of row r drains while row r + 1 computes, and is only waited for right before row r + 1 writes its first
    // output entry. Without this every CTA of a wave computes, then every CTA stores (identical work keeps them in
    // step): measured 33 us of compute + 19 us of stores = 50 us, no overlap.
    bool stage = G == 1 && pack && !(std::getenv("PGX_SPEC_STAGE") && std::atoi(std::getenv("PGX_SPEC_STAGE")) == 0);
    // measured: 0.0285 ms against 0.0280 with the L1 prefetch at row start (any position 60..100 %): off
    const bool late_ev = std::getenv("PGX_SPEC_LATE_EV") && std::atoi(std::getenv("PGX_SPEC_LATE_EV")) != 0;  // tuning knobs
    const int late_ev_pct = std::getenv("PGX_SPEC_LATE_EV_PCT") ? std::max(0, std::min(100, std::atoi(std::getenv("PGX_SPEC_LATE_EV_PCT")))) : 85;
    std::vector<int> out_col0(n_steps, -1);
    const int pitch = out_elems | 1;
    if (stage) {
        for (int g = 0; g < n_segs && stage; ++g) {
            const int32_t* sg = segs + g * SEG_WORDS;
            const int ow = seg_owner[g];
            // a segment must be exactly one step output, and no step output may feed two segments
            if (ld_i64(sg) != objs[ow].off || sg[2] != objs[ow].size || out_col0[ow] >= 0) stage = false;
            out_col0[ow] = sg[3];
        }
        for (const Step& st : steps)
            for (const Op& o : st.ops)
                if (o.work && !o.pairs.empty() && out_col0[o.owner] >= 0) stage = false;  // evidence-indexed output table
        if (!stage) std::fill(out_col0.begin(), out_col0.end(), -1);
    }
    int64_t stage_elems = 0;
    if (stage) {
        // the other shared-memory objects (evidence-indexed work tables) follow the staging block, swizzled as before
        int64_t top = 0;
        for (int si = 0; si < n_steps; ++si)
            if (in_smem[si] && out_col0[si] < 0) objs[si].new_off = top, top += objs[si].size;
        stage_elems = 32LL * pitch;
        packed_entries = top;
    }
    // Split mode (two warps per row, PGX_SPEC_SPLIT): the steps are 2-coloured along ONE cut of the junction tree — the
    // ancestors of a collect message X (the subtree below the cut edge) plus every later step that reads one of them
    // directly or reads only colour-1 values (the distribute pass and the marginals inside that subtree). Each warp
    // runs its colour with its own values in registers; only what crosses the cut goes through shared memory, with a
    // CTA barrier in front of the first level that consumes it. Any colouring is correct; X is chosen for balance.
    std::vector<char> color(n_steps, 0);
    bool split = false;
    if (G == 2 && pack && std::getenv("PGX_SPEC_SPLIT") && std::atoi(std::getenv("PGX_SPEC_SPLIT")) != 0) {
        std::vector<double> cost(n_steps);
        double total = 0;
        for (int si = 0; si < n_steps; ++si) total += cost[si] = (double)steps[si].out_size * ((double)steps[si].sum_size * steps[si].K + 2.0);
        std::vector<std::vector<char>> anc(n_steps, std::vector<char>(n_steps, 0));  // anc[s][a]: s depends on a
        for (int si = 0; si < n_steps; ++si)
            for (const Op& o : steps[si].ops)
                if (o.work && o.owner >= 0) {
                    anc[si][o.owner] = 1;
                    for (int a = 0; a < n_steps; ++a) anc[si][a] |= anc[o.owner][a];
                }
        double best = 1e300;
        std::vector<char> cand(n_steps);
        for (int x = 0; x < n_steps; ++x) {
            double c1 = 0;
            int cross = 0;
            for (int si = 0; si < n_steps; ++si) {
                bool one = si == x || anc[x][si];
                if (!one && si > x) {
                    bool direct = false, all1 = true, any = false;
                    for (const Op& o : steps[si].ops)
                        if (o.work && o.owner >= 0) {
                            any = true;
                            if (anc[x][o.owner]) direct = true;
                            if (!cand[o.owner]) all1 = false;
                        }
                    one = direct || (any && all1);
                }
                cand[si] = one;
                if (one) c1 += cost[si];
            }
            for (int si = 0; si < n_steps; ++si)
                for (const Op& o : steps[si].ops)
                    if (o.work && o.owner >= 0 && cand[o.owner] != cand[si]) ++cross;
            const double score = std::fabs(c1 - 0.5 * total) / total + 0.002 * cross;
            if (c1 > 0.2 * total && c1 < 0.8 * total && score < best) best = score, color = cand, split = true;
        }
    }
    if (split) {
        std::fill(in_smem.begin(), in_smem.end(), 0);
        for (int g = 0; g < n_segs; ++g) in_smem[seg_owner[g]] = 1;
        for (int si = 0; si < n_steps; ++si)
            for (const Op& o : steps[si].ops)
                if (o.work && o.owner >= 0 && (!o.pairs.empty() || color[o.owner] != color[si])) in_smem[o.owner] = 1;
        int64_t top = 0;
        for (int si = 0; si < n_steps; ++si)
            if (in_smem[si]) objs[si].new_off = top, top += objs[si].size;
        packed_entries = std::max<int64_t>(top, 1);
    }
    auto remap_obj = [&](int ow, int64_t e) -> int64_t { return pack ? e - objs[ow].off + objs[ow].new_off : e; };
    auto local_name = [&](int ow, int64_t e) -> std::string {
        return "m" + std::to_string(ow) + "_" + std::to_string(e - objs[ow].off);
    };
    const size_t smem = ((size_t)packed_entries * P + (size_t)stage_elems) * elem;
    if (smem > 227 * 1024 - 1024) {
        why = "work tables of one row of evidence sets exceed shared memory";
        return false;
    }
    auto wsref = [&](int64_t e) -> std::string {  // this lane's element of (remapped) entry e
        return "x" + std::to_string(e & SWZ) + "[" + std::to_string(e * P) + "]";
    };
    auto oref = [&](int ow, int64_t e) -> std::string {  // this lane's element of entry e (plan numbering) of object ow
        if (stage && ow >= 0 && out_col0[ow] >= 0) return "so[" + std::to_string(out_col0[ow] + (e - objs[ow].off)) + "]";
        return wsref(remap_obj(ow, e));
    };
    // ---- emit
    Emitter em;
    const char* T = f32 ? "float" : "double";
    em.line("// generated by pgx_spec.cu: %d steps, %lld product terms, %lld work entries (%lld before lifetime packing)", n_steps,
            (long long)terms, (long long)packed_entries, (long long)ws_entries);
    if (!pack) em.line("// work tables not packed: %s", nopack.c_str());
    em.line("typedef %s T;", T);
    em.line("#define P %d", P);
    em.line("#define FMA(a, b, c) %s((a), (b), (c))", f32 ? "fmaf" : "fma");
    // tests/hostsim/spec_host.cpp runs this very source on the CPU, one thread at a time and barrier phase by barrier
    // phase (PGX_PHASE = the phase to run; the device build runs them all)
    em.line("#ifndef PGX_PHASE");
    em.line("#define PGX_PHASE (-1)");
    em.line("#endif");
    em.line("#define PGX_WARPS %d", G);
    em.line("#define PH(p) (PGX_PHASE < 0 || PGX_PHASE == (p))");
    // a work entry that only its own thread reads: a register on the device; the CPU harness calls the kernel once per
    // barrier phase, so there it has to survive between calls
    em.line("#ifdef PGX_HOST_SIM");
    em.line("#define PGX_LOCAL(name) static T name##_a[32 * PGX_WARPS * PGX_ROWS]; T& name = name##_a[threadIdx.x];");
    em.line("#else");
    em.line("#define PGX_LOCAL(name) T name;");
    em.line("#endif");
    // reciprocal of a segment sum: MUFU.RCP64H seed + three Newton steps, no slow-path subroutine (the sums that need
    // one — subnormal — take pgx_renorm_slow; 0 -> NaN, which is what 0 * (1 / 0) has to give anyway)
    if (!f32) {
        em.line("#ifdef PGX_HOST_SIM");
        em.line("static inline double pgx_rcp(double s) { return 1.0 / s; }");
        em.line("#else");
        em.line("__device__ __forceinline__ double pgx_rcp(double s) {");
        em.line("  double r;");
        em.line("  asm(\"rcp.approx.ftz.f64 %%0, %%1;\" : \"=d\"(r) : \"d\"(s));");
        em.line("  r = fma(fma(-s, r, 1.0), r, r);");
        em.line("  r = fma(fma(-s, r, 1.0), r, r);");
        em.line("  r = fma(fma(-s, r, 1.0), r, r);");
        em.line("  return r;");
        em.line("}");
        em.line("#endif");
    } else {
        em.line("#define pgx_rcp(s) (1.0f / (s))");
    }
    // rare path of the output normalisation (a sum so small that its reciprocal could overflow): one out-of-line copy
    em.line("__device__ __noinline__ void pgx_renorm_slow(T* base, int lane, int e0, int n, T s) {");
    em.line("  for (int i = 0; i < n; ++i) { T* p = base + (e0 + i) * P + (lane ^ ((e0 + i) & %d)); *p = *p / s; }", SWZ);
    em.line("}");
    {
        std::string cm = "__device__ const int c_colmap[] = {";
        std::vector<int64_t> colmap(out_elems, 0);
        for (int g = 0; g < n_segs; ++g) {
            const int32_t* sg = segs + g * SEG_WORDS;
            for (int i = 0; i < sg[2]; ++i) colmap[sg[3] + i] = remap_obj(seg_owner[g], ld_i64(sg) + i) * P;
        }
        for (int j = 0; j < out_elems; ++j) cm += std::to_string(colmap[j]) + (j + 1 < out_elems ? "," : "");
        cm += "};";
        em.s += cm + "\n";
    }
    // resident CTAs per SM the compiler should plan for: what shared memory allows, but never so many that a thread gets
    // fewer than 255 registers when one warp runs the row (alarm: 9 CTAs x 168 registers + spills 0.080 ms, 8 x 254
    // 0.054 ms)
    // R independent rows per CTA (G == 1 only), kept in step by a CTA barrier every few steps: the straight-line code is
    // never re-executed, so the kernel is bound by instruction FETCH (ncu: no_instructions; ~1.2 instructions per clock
    // and SM whatever the occupancy) unless the warps of an SM walk through it together and share the fetched lines
    int R = 1, sync_units = 6;
    if (G == 1) {
        R = PGX_SPEC_DEFAULT_ROWS;
        if (const char* e = std::getenv("PGX_SPEC_ROWS")) R = std::max(1, std::min(16, std::atoi(e)));  // tuning knobs
        if (const char* e = std::getenv("PGX_SPEC_SYNC")) sync_units = std::max(1, std::atoi(e));
        while (R > 1 && (size_t)R * smem > 227 * 1024 - 1024) --R;
    }
    // (fp32 mode: values take one register, 16 CTAs x 128 registers measured best: 0.0217 -> 0.0165 ms on alarm)
    int min_ctas = (int)std::max<size_t>(1, std::min<size_t>(G == 1 ? std::max(1, (f32 ? 16 : 8) / R) : 32, (227 * 1024) / ((size_t)R * smem + 1024)));
    if (const char* e = std::getenv("PGX_SPEC_MINCTAS")) min_ctas = std::max(1, std::min(32, std::atoi(e)));  // tuning knob
    em.line("#define PGX_ROWS %d", R);
    em.s += "@@SEGTAB@@";
    if (stage) {
        em.line("#ifndef PGX_HOST_SIM");
        em.line("__device__ __forceinline__ void pgx_bulk_store(void* dst, const void* src, unsigned bytes) {");
        em.line("  const unsigned s = (unsigned)__cvta_generic_to_shared(src);");
        em.line("  asm volatile(\"cp.async.bulk.global.shared::cta.bulk_group [%%0], [%%1], %%2;\" :: \"l\"(dst), \"r\"(s), \"r\"(bytes) : \"memory\");");
        em.line("  asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");");
        em.line("}");
        em.line("__device__ __forceinline__ void pgx_bulk_wait() { asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\"); }");
        em.line("__device__ __forceinline__ void pgx_fence_async() { asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\"); }");
        em.line("#else");
        em.line("#define pgx_bulk_wait()");
        em.line("#endif");
    }
    em.line("extern \"C\" __global__ void __launch_bounds__(%d, %d) k_plan_spec(const T* __restrict__ cst, const int* __restrict__ ev,",
            32 * G * R, min_ctas);
    em.line("                                                            T* __restrict__ out, long long B) {");
    em.line("  extern __shared__ __align__(16) unsigned char smem_raw[];");
    em.line("  const int lane = threadIdx.x & 31;");
    if (G == 1) {
        em.line("  const int rowi = threadIdx.x >> 5;  // this warp's row inside the CTA");
        em.line("  const int warp = 0;");
    } else {
        em.line("  const int rowi = 0;");
        em.line("  const int warp = threadIdx.x >> 5;");
    }
    em.line("  T* const base = reinterpret_cast<T*>(smem_raw) + rowi * %lld;", (long long)(packed_entries * P + stage_elems));
    em.line("  T* const base2 = base + %lld;", (long long)stage_elems);
    if (stage) em.line("  T* const so = base + lane * %d;  // this evidence set's posterior row in the staging block", pitch);
    if (!stage || packed_entries > 0)
        for (int c = 0; c <= SWZ; ++c) em.line("  T* const x%d = base2 + (lane ^ %d);", c, c);
    if (stage) {
        em.line("  const long long n_rows = (B + 31) / 32;");
        // the evidence of the NEXT row is prefetched into L1 at the start of this row. Loading it into registers at row
        // start spills (the spill store then waits for the load it was meant to hide); loading it into registers late
        // in the row (PGX_SPEC_LATE_EV=1) measured no better than the prefetch
        if (late_ev && n_ev > 0) {
            em.line("  long long pb = (long long)blockIdx.x * 32 + lane;");
            em.line("  if (pb >= B) pb = B - 1;");
            for (int j = 0; j < n_ev; ++j) em.line("  int pe%d = ev[pb * %d + %d];", j, n_ev, j);
        }
        em.line("  for (long long row = blockIdx.x; row < n_rows; row += gridDim.x) {");
        em.line("  const long long row0 = row * 32;");
    } else {
        em.line("  const long long row0 = ((long long)blockIdx.x * %d + rowi) * 32;", R);
    }
    em.line("  long long b = row0 + lane;");
    em.line("  if (b >= B) b = B - 1;");
    for (int j = 0; j < n_ev; ++j) {
        if (stage && late_ev)
            em.line("  int e%d = pe%d;", j, j);
        else
            em.line("  int e%d = ev[b * %d + %d];", j, n_ev, j);
        em.line("  e%d = e%d < 0 ? 0 : (e%d > %d ? %d : e%d);", j, j, j, ev_card[j] - 1, ev_card[j] - 1, j);
    }
    if (stage && n_ev > 0 && !late_ev) {
        em.line("#ifndef PGX_HOST_SIM");
        em.line("  { long long pb = (row + gridDim.x) * 32 + lane; if (pb >= B) pb = B - 1;");
        em.line("    asm volatile(\"prefetch.global.L1 [%%0];\" :: \"l\"(ev + pb * %d)); }", n_ev);
        em.line("#endif");
    }
    // A unit = a run of consecutive output entries of one step, the grain of work handed to a warp
    struct Unit {
        std::string code;
        double cost;
        int level;
        int step;
    };
    std::vector<Unit> units;
    int n_levels = 0;
    std::vector<double> level_cost;
    for (const Step& st : steps) {
        n_levels = std::max(n_levels, st.level + 1);
        if ((int)level_cost.size() < n_levels) level_cost.resize(n_levels, 0.0);
        level_cost[st.level] += (double)st.out_size * ((double)st.sum_size * st.K + 2.0);
    }
    // registers: G warps x min_ctas CTAs share 64 K registers
    const int reg_budget = std::min(255, (65536 / (32 * G * min_ctas)) & ~7);
    const bool factor_common = !(std::getenv("PGX_SPEC_FACTOR") && std::atoi(std::getenv("PGX_SPEC_FACTOR")) == 0);  // tuning knob
    int n_acc = PGX_SPEC_DEFAULT_ACC;
    if (const char* e = std::getenv("PGX_SPEC_ACC")) n_acc = std::max(1, std::min(4, std::atoi(e)));  // tuning knob
    const bool use_select = !(std::getenv("PGX_SPEC_SELECT") && std::atoi(std::getenv("PGX_SPEC_SELECT")) == 0);  // tuning knob
    int64_t kept = 0, loads = 0, flops = 0, n_known = 0, n_locals = 0;
    std::string seg_table;  // normaliser table + slow path, spliced in front of the kernel
    std::string local_decls;
    std::map<std::pair<int, int64_t>, double> known;  // (producing step, entry) -> value, for evidence-independent entries
    // distinct elements held in registers per group of output entries
    const int cap = std::max(8, std::min(f32 ? 96 : 56, (reg_budget - 40) / (f32 ? 1 : 2)));
    for (int si = 0; si < n_steps; ++si) {
        const Step& st = steps[si];
        Emitter ue;  // code of the current unit
        Emitter pre;  // per-lane operand bases, repeated in every unit of the step
        pre.line("  { // step %d: level %d, %d operands, %lld outputs x %lld summed", si, st.level, st.K, (long long)st.out_size,
                 (long long)st.sum_size);
        int64_t cost0 = 0;
        // a step that is large for its level is cut so that the level can be spread over the G warps
        int64_t n_cut = 1;
        if (G > 1 && !split) {
            const double c = (double)st.out_size * ((double)st.sum_size * st.K + 2.0);
            n_cut = (int64_t)(c / (level_cost[st.level] / (2.0 * G)) + 0.5);
            n_cut = std::max<int64_t>(1, std::min<int64_t>(n_cut, st.out_size));
        }
        const int64_t per_unit = (st.out_size + n_cut - 1) / n_cut;
        // evidence part of the operand bases (per lane)
        std::vector<std::string> q(st.K);
        for (int k = 0; k < st.K; ++k) {
            const Op& o = st.ops[k];
            if (o.pairs.empty()) continue;
            std::string expr;
            for (auto& pr : o.pairs) {
                if (!expr.empty()) expr += " + ";
                expr += "e" + std::to_string(pr.first) + " * " + std::to_string(pr.second);
            }
            q[k] = "q" + std::to_string(k);
            if (o.work)
                pre.line("    const int %s = (%s);", q[k].c_str(), expr.c_str());
            else
                pre.line("    const T* const %s = cst + (%s);", q[k].c_str(), expr.c_str());
        }
        // element index of every (o, s, k) by mixed-radix counting
        std::vector<int> od(std::max(st.A, 1), 0);
        struct Term {
            double coef;
            std::vector<std::string> f;
        };
        std::vector<std::vector<Term>> entry_terms;  // of the current group
        std::vector<Term> entry_den;                 // divisor of each entry
        std::vector<char> entry_has_den;
        std::vector<std::string> entry_out;  // where the entry goes: shared-memory reference or local variable
        std::map<std::string, std::string> group_loads;  // name -> load expression
        auto flush = [&]() {
            if (entry_terms.empty()) return;
            ue.line("    {");
            for (auto& kv : group_loads) ue.line("      const T %s = %s;", kv.first.c_str(), kv.second.c_str());
            loads += (int64_t)group_loads.size();
            for (size_t i = 0; i < entry_terms.size(); ++i) {
                std::string acc;
                double konst = 0.0;
                bool have_konst = false;
                std::string body;
                bool first = true;
                // n_acc > 1: the terms of an entry go round robin onto n_acc partial sums added at the end, which cuts the
                // dependent fma chain of a long sum (the order of the additions changes: ~1e-16 relative)
                // a factor every term of the entry shares (an operand that does not depend on the summed variables) is
                // taken out of the sum: (sum of the reduced terms) * factor
                std::vector<std::string> common;
                if (factor_common && entry_terms[i].size() >= 2) {
                    bool all_dyn = true;
                    for (const Term& t : entry_terms[i]) all_dyn = all_dyn && !t.f.empty();
                    if (all_dyn) {
                        common = entry_terms[i][0].f;
                        for (size_t ti2 = 1; ti2 < entry_terms[i].size() && !common.empty(); ++ti2) {
                            std::vector<std::string> rest = entry_terms[i][ti2].f, keep;
                            for (const std::string& f : common) {
                                auto it = std::find(rest.begin(), rest.end(), f);
                                if (it != rest.end()) {
                                    keep.push_back(f);
                                    rest.erase(it);
                                }
                            }
                            common = keep;
                        }
                        for (Term& t : entry_terms[i])
                            for (const std::string& f : common) t.f.erase(std::find(t.f.begin(), t.f.end(), f));
                    }
                }
                int n_nc = 0;
                for (const Term& t : entry_terms[i]) n_nc += !t.f.empty();
                const int na = (n_acc > 1 && n_nc >= 2 * n_acc) ? n_acc : 1;
                std::vector<char> firsts(na, 1);
                int ti = 0;
                for (const Term& t : entry_terms[i]) {
                    if (t.f.empty()) {
                        konst += t.coef;
                        have_konst = true;
                        continue;
                    }
                    const std::string a = na > 1 ? "a" + std::to_string(ti % na) : std::string("a");
                    first = firsts[ti % na];
                    firsts[ti % na] = 0;
                    ++ti;
                    std::string prod = t.f[0];
                    for (size_t j = 1; j + 1 < t.f.size(); ++j) prod += " * " + t.f[j], ++flops;
                    const std::string c = lit(t.coef, f32);
                    std::string expr;
                    if (t.f.size() == 1) {
                        if (first)
                            expr = t.coef == 1.0 ? prod : prod + " * " + c;
                        else
                            expr = t.coef == 1.0 ? a + " + " + prod : "FMA(" + prod + ", " + c + ", " + a + ")";
                    } else {
                        const std::string& last = t.f.back();
                        if (t.coef == 1.0) {
                            expr = first ? prod + " * " + last : "FMA(" + prod + ", " + last + ", " + a + ")";
                        } else {
                            ++flops;
                            expr = first ? "(" + prod + " * " + last + ") * " + c : "FMA(" + prod + " * " + last + ", " + c + ", " + a + ")";
                        }
                    }
                    ++flops;
                    body += std::string("      ") + (first ? "T " + a + " = " : a + " = ") + expr + ";\n";
                    first = false;
                }
                if (na > 1) {
                    std::string sum = "a0";
                    for (int j = 1; j < na; ++j) sum += " + a" + std::to_string(j), ++flops;
                    body += "      T a = " + sum + ";\n";
                }
                first = n_nc == 0;
                if (first) {
                    body = "      T a = " + lit(have_konst ? konst : 0.0, f32) + ";\n";
                } else if (have_konst && konst != 0.0) {
                    body += "      a += " + lit(konst, f32) + ";\n";
                    ++flops;
                }
                for (const std::string& f : common) body += "      a = a * " + f + ";\n", ++flops;
                if (entry_has_den[i]) {
                    const Term& d = entry_den[i];
                    std::string de;
                    for (const std::string& f : d.f) de += (de.empty() ? "" : " * ") + f, ++flops;
                    if (de.empty() || d.coef != 1.0) de += (de.empty() ? "" : " * ") + lit(d.coef, f32);
                    body += "      a = a / (" + de + ");\n";
                    body += "      a = (a != a) ? (T)0 : a;\n";  // 0 / 0 -> 0
                    flops += 12;
                }
                ue.s += "      {\n" + body;
                ue.line("      %s = a; }", entry_out[i].c_str());
            }
            ue.line("    }");
            entry_terms.clear();
            entry_den.clear();
            entry_has_den.clear();
            entry_out.clear();
            group_loads.clear();
        };
        auto close_unit = [&]() {
            flush();
            if (ue.s.empty()) return;
            units.push_back({pre.s + ue.s + "  }\n", (double)(flops + loads - cost0) + 4.0, st.level, si});
            ue.s.clear();
            cost0 = flops + loads;
        };
        cost0 = flops + loads;
        for (int64_t o = 0; o < st.out_size; ++o) {
            if (o > 0 && o % per_unit == 0) close_unit();
            std::vector<int64_t> ob(st.K);
            for (int k = 0; k < st.K; ++k) {
                int64_t e = st.ops[k].base;
                for (int a = 0; a < st.A; ++a) e += (int64_t)od[a] * st.ops[k].ostr[a];
                ob[k] = e;
            }
            std::vector<Term> tl;
            std::map<std::string, std::string> need;
            Term cur_den{1.0, {}};
            bool cur_has_den = false;
            auto add_factor = [&](int k, int64_t e, Term& t) {
                const Op& op = st.ops[k];
                    if (!op.work && op.pairs.empty()) {
                        t.coef *= cval(e);
                    } else if (op.work && op.pairs.empty() && pack && known.count({op.owner, e})) {
                        t.coef *= known[{op.owner, e}];  // evidence-independent entry: evaluated here, on the host
                    } else if (op.work && op.pairs.empty() && pack && !in_smem[op.owner]) {
                        t.f.push_back(local_name(op.owner, e));  // a value of this thread: no load
                    } else if (op.work && op.pairs.empty()) {
                        const std::string name = "w" + std::to_string(op.owner) + "_" + std::to_string(e - (pack ? objs[op.owner].off : 0));
                        need[name] = oref(op.owner, e);
                        t.f.push_back(name);
                    } else if (op.work) {
                        // evidence-indexed work table: the whole table is one object, so the remap is a constant shift
                        const int64_t ne = remap_obj(op.owner, e);
                        const std::string name = "v" + std::to_string(k) + "_" + std::to_string(ne);
                        const std::string ed = "(" + q[k] + " + " + std::to_string(ne) + ")";
                        need[name] = "base2[" + ed + " * P + (lane ^ (" + ed + " & " + std::to_string(SWZ) + "))]";
                        t.f.push_back(name);
                    } else if (use_select && op.pairs.size() == 1 && ev_card[op.pairs[0].first] <= 4) {
                        // CPT entry indexed by ONE observed variable with few states: a select among immediates instead
                        // of a gather (the load/store path is the busy one, the ALU is not)
                        const int slot = op.pairs[0].first, card = ev_card[slot];
                        bool same = true;
                        for (int c = 1; c < card; ++c) same = same && cval(e + (int64_t)c * op.pairs[0].second) == cval(e);
                        if (same) {
                            t.coef *= cval(e);
                        } else {
                            const std::string name = "g" + std::to_string(k) + "_" + std::to_string(e);
                            std::string expr = lit(cval(e + (int64_t)(card - 1) * op.pairs[0].second), f32);
                            for (int c = card - 2; c >= 0; --c)
                                expr = "(e" + std::to_string(slot) + " == " + std::to_string(c) + " ? " +
                                       lit(cval(e + (int64_t)c * op.pairs[0].second), f32) + " : " + expr + ")";
                            need[name] = expr;
                            t.f.push_back(name);
                        }
                    } else {
                        const std::string name = "g" + std::to_string(k) + "_" + std::to_string(e);
                        need[name] = "__ldg(" + q[k] + " + " + std::to_string(e) + ")";
                        t.f.push_back(name);
                    }
                            };

            std::vector<int> sd(std::max(st.S, 1), 0);
            for (int64_t s = 0; s < st.sum_size; ++s) {
                Term t;
                t.coef = 1.0;
                for (int k = 0; k < st.K; ++k) {
                    const Op& op = st.ops[k];
                    int64_t e = ob[k];
                    for (int a = 0; a < st.S; ++a) e += (int64_t)sd[a] * op.sstr[a];
                    if (op.div) continue;
                    add_factor(k, e, t);
                }
                ++stats.terms;
                if (t.coef != 0.0) {
                    tl.push_back(std::move(t));
                    ++kept;
                }
                for (int a = st.S - 1; a >= 0; --a) {
                    if (++sd[a] < st.sdims[a]) break;
                    sd[a] = 0;
                }
            }
            // an entry without evidence-dependent factors is a constant of the plan (messages out of subtrees that hold
            // no observed variable): later steps fold it into their coefficients
            {
                bool all_const = true;
                double v = 0.0;
                for (const Term& t : tl) {
                    if (!t.f.empty()) {
                        all_const = false;
                        break;
                    }
                    v += t.coef;
                }
                Term den;
                den.coef = 1.0;
                bool has_den = false;
                for (int k = 0; k < st.K; ++k)
                    if (st.ops[k].div) add_factor(k, ob[k], den), has_den = true;
                if (has_den && den.f.empty()) {
                    // a constant divisor: 0 / 0 -> 0, x / 0 -> inf, like the reference's divide (DiscreteFactor.py:859-863)
                    if (all_const) {
                        v = (v == 0.0 && den.coef == 0.0) ? 0.0 : v / den.coef;
                        tl.clear();
                        Term c;
                        c.coef = v;
                        tl.push_back(c);
                        has_den = false;
                    }
                } else if (has_den) {
                    all_const = false;
                }
                if (!has_den) den = Term{1.0, {}};
                cur_den = den;
                cur_has_den = has_den;
                if (all_const && pack) known[{si, st.out_off + o}] = f32 ? (double)(float)v : v, ++n_known;
            }
            // only the elements of kept terms are loaded
            std::map<std::string, std::string> used;
            for (const Term& t : tl)
                for (const std::string& f : t.f)
                    if (need.count(f)) used[f] = need[f];
            for (const std::string& f : cur_den.f)
                if (need.count(f)) used[f] = need[f];
            size_t merged = group_loads.size();
            for (auto& kv : used)
                if (!group_loads.count(kv.first)) ++merged;
            if (merged > (size_t)cap && !entry_terms.empty()) flush();
            for (auto& kv : used) group_loads[kv.first] = kv.second;
            entry_terms.push_back(std::move(tl));
            if (!cur_has_den) cur_den.coef = 1.0, cur_den.f.clear();
            entry_den.push_back(cur_has_den ? cur_den : Term{1.0, {}});
            entry_has_den.push_back(cur_has_den);
            if (in_smem[si]) {
                entry_out.push_back(oref(si, st.out_off + o));
            } else {
                entry_out.push_back(local_name(si, st.out_off + o));
                local_decls += "  PGX_LOCAL(" + local_name(si, st.out_off + o) + ")\n";
                ++n_locals;
            }
            for (int a = st.A - 1; a >= 0; --a) {
                if (++od[a] < st.odims[a]) break;
                od[a] = 0;
            }
        }
        close_unit();
    }
    // ---- normalise the output segments in place (lane = evidence set)
    if (stage) {
        // ONE branch per row instead of one per variable: the sums of all segments first, then either the straight-line
        // reciprocal path or (some sum below 1e-30, zero or NaN: impossible evidence, underflow) an out-of-line loop that
        // divides. With a branch per segment the instruction after each branch waited for its cache line (ncu source
        // page: 32 MUFU.RCP64H held 18 % of all stall samples, all of them no_instructions; BSYNC another 9 %).
        Emitter ue;
        std::string tab;
        int n_norm = 0;
        ue.line("  { // normalise");
        for (int g = 0; g < n_segs; ++g) {
            const int32_t* sg = segs + g * SEG_WORDS;
            if (!(sg[4] & SEG_NORMALIZE)) continue;
            const int c0 = out_col0[seg_owner[g]], n = sg[2];
            std::string sum = "so[" + std::to_string(c0) + "]";
            for (int i = 1; i < n; ++i) sum += " + so[" + std::to_string(c0 + i) + "]";
            ue.line("    const T s%d = %s;", g, sum.c_str());
            tab += std::to_string(c0) + "," + std::to_string(n) + ",";
            ++n_norm;
        }
        std::string ok = "true";
        for (int g = 0; g < n_segs; ++g)
            if (segs[g * SEG_WORDS + 4] & SEG_NORMALIZE) ok += " & (s" + std::to_string(g) + " >= (T)1e-30)";
        ue.line("    if (%s) {", ok.c_str());
        for (int g = 0; g < n_segs; ++g) {
            const int32_t* sg = segs + g * SEG_WORDS;
            if (!(sg[4] & SEG_NORMALIZE)) continue;
            const int c0 = out_col0[seg_owner[g]], n = sg[2];
            ue.line("      { const T r = pgx_rcp(s%d);", g);
            for (int i = 0; i < n; ++i) ue.line("        so[%d] = so[%d] * r;", c0 + i, c0 + i);
            ue.line("      }");
            flops += 2 * n + 8;
            loads += n;
        }
        ue.line("    } else {");
        ue.line("      pgx_renorm_all_slow(so);");
        ue.line("    }");
        ue.line("  }");
        if (n_norm > 0) {
            seg_table = "__device__ const int c_segtab[] = {" + tab + "0};\n";
            seg_table += "__device__ __noinline__ void pgx_renorm_all_slow(T* so) {\n";
            seg_table += "  for (int g = 0; g < " + std::to_string(n_norm) + "; ++g) {\n";
            seg_table += "    const int c0 = c_segtab[2 * g], n = c_segtab[2 * g + 1];\n";
            seg_table += "    T s = (T)0;\n    for (int i = 0; i < n; ++i) s += so[c0 + i];\n";
            seg_table += "    for (int i = 0; i < n; ++i) so[c0 + i] = so[c0 + i] / s;\n  }\n}\n";
            units.push_back({ue.s, 400.0, n_levels, n_steps});
        }
    }
    for (int g = 0; g < n_segs && !stage; ++g) {
        const int32_t* sg = segs + g * SEG_WORDS;
        if (!(sg[4] & SEG_NORMALIZE)) continue;
        const int64_t off = remap_obj(seg_owner[g], ld_i64(sg));
        const int64_t off0 = ld_i64(sg);
        const int sow = seg_owner[g];
        const int n = sg[2];
        Emitter ue;
        ue.line("  { // segment %d", g);
        for (int i = 0; i < n; ++i) ue.line("    const T v%d = %s;", i, oref(sow, off0 + i).c_str());
        std::string sum = "v0";
        for (int i = 1; i < n; ++i) sum += " + v" + std::to_string(i);
        ue.line("    const T s = %s;", sum.c_str());
        // values / values.sum() as values * (1 / sum): one division per segment (<= 1 ulp from the quotient);
        // 0 * inf = NaN like 0 / 0; where 1 / sum could overflow, divide
        ue.line("    if (s != (T)0 && (s < (T)0 ? -s : s) < (T)1e-30) {");
        if (stage)
            ue.line("      for (int i = 0; i < %d; ++i) so[%d + i] = so[%d + i] / s;", n, out_col0[sow], out_col0[sow]);
        else
            ue.line("      pgx_renorm_slow(base2, lane, %lld, %d, s);", (long long)off, n);
        ue.line("    } else {");
        ue.line("      const T r = pgx_rcp(s);");
        for (int i = 0; i < n; ++i) ue.line("      %s = v%d * r;", oref(sow, off0 + i).c_str(), i);
        ue.line("    }");
        ue.line("  }");
        flops += 2 * n + 8;
        loads += n;
        units.push_back({ue.s, 3.0 * n + 30.0, n_levels, n_steps + g});
    }
    // ---- the levels: units dealt to the warps of the row, largest first onto the least loaded warp; one CTA barrier
    // between levels (none at all when one warp runs the row: lane l only reads what lane l wrote)
    int phase = 0;
    if (G == 1) {
        em.line("  if (PH(0)) {");
        em.s += local_decls;
        // plan order (level by level). Emitting every step right before its first consumer (depth first from the
        // outputs) was tried to shorten live ranges: ptxas then spills MORE (448 B of stack at 255 registers against
        // none) — the collect messages have to stay alive until the distribute pass whatever the order is
        int ui = 0;
        const bool skip_compute = std::getenv("PGX_SPEC_DEBUG_SKIP_COMPUTE") != nullptr;  // timing experiments only
        bool waited = !stage;
        // PGX_SPEC_DEFER=1: steps that only feed the output (marginals) emitted last, so that the previous row's bulk store
        // — all CTAs issue theirs at about the same time, a 27 MB burst that takes ~5 us to drain — has the whole row
        // to finish before the staging block is written again (ncu: the wait in front of the first output write holds
        // 5 % of the stall samples). Measured: 0.0393 ms against 0.0300 — the marginals' operands then live to the
        // end of the row and spill (592 B of stack). Off.
        std::vector<Unit> ordered;
        if (stage && std::getenv("PGX_SPEC_DEFER") && std::atoi(std::getenv("PGX_SPEC_DEFER")) != 0) {  // measured slower: off
            std::vector<char> consumed(n_steps, 0);
            for (const Step& st2 : steps)
                for (const Op& o : st2.ops)
                    if (o.work && o.owner >= 0) consumed[o.owner] = 1;
            for (const Unit& u : units)
                if (u.step < n_steps && !(in_smem[u.step] && !consumed[u.step])) ordered.push_back(u);
            for (const Unit& u : units)
                if (u.step < n_steps && in_smem[u.step] && !consumed[u.step]) ordered.push_back(u);
            for (const Unit& u : units)
                if (u.step >= n_steps) ordered.push_back(u);
        } else {
            ordered = units;
        }
        size_t uidx = 0;
        const size_t ev_at = ordered.size() * (size_t)late_ev_pct / 100;
        bool ev_loaded = !(stage && late_ev && n_ev > 0);
        auto load_next_ev = [&]() {
            em.line("  pb = (row + gridDim.x) * 32 + lane;");
            em.line("  if (pb >= B) pb = B - 1;");
            for (int j = 0; j < n_ev; ++j) em.line("  pe%d = ev[pb * %d + %d];", j, n_ev, j);
            ev_loaded = true;
        };
        for (const Unit& u : ordered) {
            if (!ev_loaded && uidx++ >= ev_at) load_next_ev();
            if (skip_compute) break;
            if (!waited && (u.step >= n_steps || in_smem[u.step])) {
                // the previous row's bulk store reads the staging block: it must have finished before it is rewritten
                if (!std::getenv("PGX_SPEC_DEBUG_NOWAIT")) {  // timing experiments only
                    em.line("  if (lane == 0) pgx_bulk_wait();");
                    em.line("  __syncwarp();");
                }
                waited = true;
            }
            em.s += u.code;
            if (R > 1 && ++ui % sync_units == 0) em.line("  __syncthreads();");
        }
        if (!ev_loaded) load_next_ev();
        em.line("  }");
        phase = 1;
    }
    if (split) {
        em.s += local_decls;
        // phases: a barrier goes in front of the first level that reads a value the OTHER warp produced since the last one
        auto ucolor = [&](const Unit& u) { return u.step < n_steps ? color[u.step] : color[seg_owner[u.step - n_steps]]; };
        std::vector<int> phase_of_level(n_levels + 1, 0);
        int cur_phase = 0, phase_start = 0;
        for (int lv = 0; lv <= n_levels; ++lv) {
            bool need = false;
            for (int si = 0; si < n_steps && !need; ++si)
                if (steps[si].level == lv)
                    for (const Op& o : steps[si].ops)
                        if (o.work && o.owner >= 0 && color[o.owner] != color[si] && steps[o.owner].level >= phase_start) need = true;
            if (need) ++cur_phase, phase_start = lv;
            phase_of_level[lv] = cur_phase;
        }
        for (int ph = 0; ph <= cur_phase; ++ph) {
            em.line("  if (PH(%d)) {", ph);
            for (int c = 0; c < 2; ++c) {
                em.line(c == 0 ? "  if (warp == 0) {" : "  } else {");
                for (const Unit& u : units)
                    if (ucolor(u) == c && phase_of_level[u.level] == ph) em.s += u.code;
            }
            em.line("  }");
            em.line("  }");
            em.line("  __syncthreads();");
        }
        phase = cur_phase + 1;
    }
    for (int lv = 0; lv <= n_levels && G > 1 && !split; ++lv) {
        std::vector<int> idx;
        for (int u = 0; u < (int)units.size(); ++u)
            if (units[u].level == lv) idx.push_back(u);
        if (idx.empty()) continue;
        em.line("  if (PH(%d)) { // level %d", phase, lv);
        if (G == 1) {
            for (int u : idx) em.s += units[u].code;
        } else {
            std::vector<int> order = idx;
            std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return units[a].cost > units[c].cost; });
            std::vector<double> load(G, 0.0);
            std::vector<std::vector<int>> mine(G);
            for (int u : order) {
                const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
                load[w] += units[u].cost;
                mine[w].push_back(u);
            }
            em.line("  switch (warp) {");
            for (int w = 0; w < G; ++w) {
                if (mine[w].empty()) continue;
                std::sort(mine[w].begin(), mine[w].end());  // program order inside a warp
                em.line("  case %d: {", w);
                for (int u : mine[w]) em.s += units[u].code;
                em.line("  } break;");
            }
            em.line("  default: break;");
            em.line("  }");
        }
        em.line("  }");
        if (G > 1) em.line("  __syncthreads();");
        ++phase;
    }
    // ---- output stage: the CTA's 32 rows of out[] are one contiguous range; lanes = columns (coalesced stores), the
    // warps of the row take every G-th row
    em.line("  // %lld work entries were evidence independent (evaluated by the generator)", (long long)n_known);
    if (G == 1) em.line("  __syncwarp();");
    em.line("  if (PH(%d)) {", phase);
    if (std::getenv("PGX_SPEC_DEBUG_SKIP_OUTPUT")) em.line(stage ? "  if (B > 0) continue;" : "  if (B > 0) return;");  // timing experiments only
    em.line("  const int rows = (int)((B - row0) < 32 ? (B - row0) : 32);");
    em.line("  T* const dst = out + row0 * %d;", out_elems);
    if (stage) {
        em.line("#ifndef PGX_HOST_SIM");
        em.line("  if (%d && rows == 32 && (((unsigned long long)dst) & 15) == 0) {", pitch == out_elems ? 1 : 0);
        em.line("    pgx_fence_async();  // this lane's writes to the staging block -> visible to the async proxy");
        em.line("    __syncwarp();");
        em.line("    if (lane == 0) pgx_bulk_store(dst, base, %du);", (unsigned)(32 * out_elems * elem));
        em.line("  } else");
        em.line("#endif");
        em.line("  {");
        em.line("    for (int j = lane; j < %d; j += 32)", out_elems);
        em.line("      for (int bb = 0; bb < rows; ++bb) dst[(long long)bb * %d + j] = base[bb * %d + j];", out_elems, pitch);
        em.line("    __syncwarp();");
        em.line("  }");
        em.line("  }");   // PH
        em.line("  }");   // row loop
        em.line("  if (lane == 0) pgx_bulk_wait();  // shared memory must outlive the last bulk store");
        em.line("}");
        em.line("#define PGX_N_PHASES %d", phase + 1);
    } else {
    em.line("#pragma unroll 1");
    em.line("  for (int j = lane; j < %d; j += 32) {", out_elems);
    em.line("    const T* const src = base + c_colmap[j];");
    em.line("    const int sw = (c_colmap[j] >> 5) & %d;", SWZ);
    em.line("    T* const d = dst + j;");
    em.line("    if (rows == 32) {");
    em.line("#pragma unroll");
    em.line("      for (int bb = warp; bb < 32; bb += %d) d[(long long)bb * %d] = src[bb ^ sw];", G, out_elems);
    em.line("    } else {");
    em.line("      for (int bb = warp; bb < rows; bb += %d) d[(long long)bb * %d] = src[bb ^ sw];", G, out_elems);
    em.line("    }");
    em.line("  }");
    em.line("  }");
    em.line("}");
    em.line("#define PGX_N_PHASES %d", phase + 1);
    }
    {
        const size_t at = em.s.find("@@SEGTAB@@");
        if (at != std::string::npos) em.s.replace(at, 10, seg_table);
    }
    source = std::move(em.s);
    stats.terms = terms;
    stats.terms_kept = kept;
    stats.loads = loads;
    stats.flops = flops;
    stats.ws_entries = packed_entries;
    stats.smem_bytes = (int64_t)smem * R;
    stats.warps = G;
    stats.rows = R;
    stats.persistent = stage ? std::max(1, min_ctas) : 0;
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// NVRTC through dlopen (the library must load on machines without the toolkit; specialisation is then unavailable)
namespace {
struct Nvrtc {
    void* h = nullptr;
    int (*create)(void**, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    int (*compile)(void*, int, const char* const*) = nullptr;
    int (*log_size)(void*, size_t*) = nullptr;
    int (*get_log)(void*, char*) = nullptr;
    int (*cubin_size)(void*, size_t*) = nullptr;
    int (*get_cubin)(void*, char*) = nullptr;
    int (*destroy)(void**) = nullptr;
    bool ok = false;
};
Nvrtc& nvrtc() {
    static Nvrtc n;
    static bool tried = false;
    if (tried) return n;
    tried = true;
    const char* names[] = {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so"};
    for (const char* nm : names) {
        n.h = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
        if (n.h) break;
    }
    if (!n.h) return n;
    n.create = (decltype(n.create))dlsym(n.h, "nvrtcCreateProgram");
    n.compile = (decltype(n.compile))dlsym(n.h, "nvrtcCompileProgram");
    n.log_size = (decltype(n.log_size))dlsym(n.h, "nvrtcGetProgramLogSize");
    n.get_log = (decltype(n.get_log))dlsym(n.h, "nvrtcGetProgramLog");
    n.cubin_size = (decltype(n.cubin_size))dlsym(n.h, "nvrtcGetCUBINSize");
    n.get_cubin = (decltype(n.get_cubin))dlsym(n.h, "nvrtcGetCUBIN");
    n.destroy = (decltype(n.destroy))dlsym(n.h, "nvrtcDestroyProgram");
    n.ok = n.create && n.compile && n.log_size && n.get_log && n.cubin_size && n.get_cubin && n.destroy;
    return n;
}
}  // namespace

// Optional on-disk cache of compiled kernels (PGX_SPEC_CACHE_DIR; off when unset): the generated source is a pure
// function of the plan and the table values, so its 64-bit FNV-1a hash (+ length) names the cubin. A second process
// serving the same model and evidence signature then pays a file read instead of ~0.5-3 s of NVRTC.
static std::string cache_path(const std::string& source) {
    const char* dir = std::getenv("PGX_SPEC_CACHE_DIR");
    if (!dir || !*dir) return "";
    uint64_t h = 1469598103934665603ull;
    for (unsigned char c : source) h = (h ^ c) * 1099511628211ull;
    char name[96];
    snprintf(name, sizeof name, "/pgx_spec_sm100a_%016llx_%zu.cubin", (unsigned long long)h, source.size());
    return std::string(dir) + name;
}

bool pgx_spec_compile(const std::string& source, std::string& cubin, std::string& log) {
    const std::string cached = cache_path(source);
    if (!cached.empty()) {
        if (FILE* f = fopen(cached.c_str(), "rb")) {
            fseek(f, 0, SEEK_END);
            const long sz = ftell(f);
            fseek(f, 0, SEEK_SET);
            cubin.resize(sz > 0 ? (size_t)sz : 0);
            const bool ok = sz > 4 && fread(&cubin[0], 1, (size_t)sz, f) == (size_t)sz && cubin.compare(0, 4, "\x7f" "ELF") == 0;
            fclose(f);
            if (ok) {
                log = "cache hit: " + cached;
                return true;
            }
        }
    }
    Nvrtc& n = nvrtc();
    if (!n.ok) {
        log = "libnvrtc not found";
        return false;
    }
    void* prog = nullptr;
    if (n.create(&prog, source.c_str(), "pgx_plan_spec.cu", 0, nullptr, nullptr) != 0) {
        log = "nvrtcCreateProgram failed";
        return false;
    }
    const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    const int rc = n.compile(prog, 3, opts);
    size_t ls = 0;
    n.log_size(prog, &ls);
    if (ls > 1) {
        log.resize(ls);
        n.get_log(prog, &log[0]);
    }
    if (rc != 0) {
        n.destroy(&prog);
        if (log.empty()) log = "nvrtcCompileProgram failed";
        return false;
    }
    size_t cs = 0;
    n.cubin_size(prog, &cs);
    cubin.resize(cs);
    n.get_cubin(prog, &cubin[0]);
    n.destroy(&prog);
    if (cs > 0 && !cached.empty()) {
        const std::string tmp = cached + ".tmp" + std::to_string((long long)getpid());
        if (FILE* f = fopen(tmp.c_str(), "wb")) {  // write + rename: concurrent ranks never see a partial file
            const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
            fclose(f);
            if (!ok || rename(tmp.c_str(), cached.c_str()) != 0) remove(tmp.c_str());
        }
    }
    return cs > 0;
}

// ------------------------------------------------------------------------------------------------------------------
// driver entry points through the runtime (no link against libcuda)
namespace {
struct Drv {
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*moduleUnload)(CUmodule) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*funcGetAttribute)(int*, CUfunction_attribute, CUfunction) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                             void**) = nullptr;
    bool ok = false;
};
Drv& drv() {
    static Drv d;
    static bool tried = false;
    if (tried) return d;
    tried = true;
    auto get = [&](const char* name, void** fn) {
        cudaDriverEntryPointQueryResult q;
        return cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) == cudaSuccess && *fn != nullptr;
    };
    d.ok = get("cuModuleLoadData", (void**)&d.moduleLoadData) && get("cuModuleGetFunction", (void**)&d.moduleGetFunction) &&
           get("cuModuleUnload", (void**)&d.moduleUnload) && get("cuFuncSetAttribute", (void**)&d.funcSetAttribute) &&
           get("cuFuncGetAttribute", (void**)&d.funcGetAttribute) && get("cuLaunchKernel", (void**)&d.launchKernel);
    return d;
}
}  // namespace

struct SpecKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int sm_count = 148;
    SpecStats stats;
};

SpecKernel* pgx_spec_build(const int32_t* pool, int64_t pool_words, const void* host_blob, int64_t table_entries, int dtype,
                           std::string& why, int warps) {
    std::string src, cubin, log;
    SpecStats stats;
    const auto t0 = std::chrono::steady_clock::now();
    if (!pgx_spec_generate(pool, pool_words, host_blob, table_entries, dtype, src, stats, why, warps)) return nullptr;
    if (!pgx_spec_compile(src, cubin, log)) {
        why = "NVRTC: " + log;
        return nullptr;
    }
    cudaFree(nullptr);  // make sure the primary context is current
    Drv& d = drv();
    if (!d.ok) {
        why = "driver entry points unavailable";
        return nullptr;
    }
    SpecKernel* k = new SpecKernel();
    if (d.moduleLoadData(&k->mod, cubin.data()) != CUDA_SUCCESS || d.moduleGetFunction(&k->fn, k->mod, "k_plan_spec") != CUDA_SUCCESS) {
        why = "cuModuleLoadData failed";
        if (k->mod) d.moduleUnload(k->mod);
        delete k;
        return nullptr;
    }
    if (d.funcSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)stats.smem_bytes) != CUDA_SUCCESS) {
        why = "cuFuncSetAttribute(shared memory) failed";
        d.moduleUnload(k->mod);
        delete k;
        return nullptr;
    }
    d.funcGetAttribute(&stats.regs, CU_FUNC_ATTRIBUTE_NUM_REGS, k->fn);
    {
        int dev = 0, sms = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0)
            k->sm_count = sms;
    }
    stats.compile_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    k->stats = stats;
    return k;
}

const SpecStats& pgx_spec_stats(const SpecKernel* k) { return k->stats; }

int pgx_spec_launch(SpecKernel* k, const void* cst, const int32_t* ev, void* out, int64_t B, void* stream, std::string& err) {
    if (B <= 0) return 0;
    long long b = B;
    void* args[] = {(void*)&cst, (void*)&ev, (void*)&out, (void*)&b};
    const int64_t per_cta = 32 * (int64_t)k->stats.rows;
    unsigned grid = (unsigned)((B + per_cta - 1) / per_cta);
    if (k->stats.persistent > 0) grid = std::min<unsigned>(grid, (unsigned)(k->sm_count * k->stats.persistent));
    const CUresult rc = drv().launchKernel(k->fn, grid, 1, 1, 32u * (unsigned)(k->stats.warps * k->stats.rows), 1, 1, (unsigned)k->stats.smem_bytes, (CUstream)stream, args, nullptr);
    if (rc != CUDA_SUCCESS) {
        err = "cuLaunchKernel(k_plan_spec) failed: " + std::to_string((int)rc);
        return -1;
    }
    return 0;
}

void pgx_spec_destroy(SpecKernel* k) {
    if (!k) return;
    if (k->mod) drv().moduleUnload(k->mod);
    delete k;
}

}  // namespace pgx

// ---- C-ABI: host-only inspection entry (declared in include/pgx.h) ------------------------------------------------
#include "../../include/pgx.h"

static int spec_default_warps() {
    if (const char* e = std::getenv("PGX_SPEC_WARPS")) return std::max(1, std::atoi(e));  // tuning knob
    return PGX_SPEC_DEFAULT_WARPS;
}

extern "C" int64_t pgx_spec_source(const pgx_plan_desc* desc, int32_t compile, char* buf, int64_t cap, int64_t* stats8) {
    if (!desc || !desc->pool || desc->pool_words < pgx::HEADER_WORDS || !desc->table_blob) return -1;
    std::string src, why;
    pgx::SpecStats st;
    if (!pgx::pgx_spec_generate(desc->pool, desc->pool_words, desc->table_blob, desc->table_entries, desc->dtype, src, st, why,
                                spec_default_warps())) {
        if (buf && cap > 0) snprintf(buf, (size_t)cap, "%s", why.c_str());
        return -2;
    }
    int64_t cubin_bytes = 0;
    if (compile) {
        std::string cubin, log;
        const auto t0 = std::chrono::steady_clock::now();
        if (!pgx::pgx_spec_compile(src, cubin, log)) {
            if (buf && cap > 0) snprintf(buf, (size_t)cap, "%s", log.c_str());
            return -3;
        }
        st.compile_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        cubin_bytes = (int64_t)cubin.size();
        if (compile == 2) src = cubin;  // hand back the cubin instead of the source (SASS inspection)
    }
    if (stats8) {
        stats8[0] = st.terms, stats8[1] = st.terms_kept, stats8[2] = st.loads, stats8[3] = st.flops;
        stats8[4] = st.ws_entries, stats8[5] = st.smem_bytes, stats8[6] = (int64_t)(st.compile_s * 1e3), stats8[7] = cubin_bytes;
    }
    if (buf && cap > 0) {
        const size_t n = std::min<size_t>(src.size(), (size_t)cap);
        memcpy(buf, src.data(), n);
    }
    return (int64_t)src.size();
}
