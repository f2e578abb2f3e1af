// pgx_fused.cuh — table-driven whole-plan kernel (K5 for small models) and its host-side "microprogram" builder.
//
// Why: the generic element function (pgx_step.cuh) spends ~50 instructions per operand load on mixed-radix
// index arithmetic (ncu, profiles/r01_v1_*). All of that arithmetic is batch invariant, so pgx_plan_create
// expands every step ONCE into two small offset tables per operand,
//
//     entry_k(o, s) = otab[o][k] + stab[s][k]            (table base folded into otab)
//
// and the kernel only adds, loads and multiplies. Work tables live in SHARED memory as [entry][32 lanes]
// (lane = evidence set), so messages and beliefs never touch HBM; HBM sees the evidence rows in and the posterior
// rows out. Steps are grouped into dependency levels (plan.py); within a level the (step, output entry) work items
// are dealt round-robin to the warps of the CTA, one __syncthreads() per level.
//
// Microprogram layout (int32 words, device global memory, read through L1 with warp-uniform addresses):
//   header[8]: 0 n_levels | 1 n_items | 2 levels_off | 3 items_off | 4 n_segs | 5 segs_off | 6 colmap_off | 7 out_elems
//   levels[n_levels+1]: first item of each level
//   items[n_items][3]:  (step record offset, first output entry, number of consecutive entries)  — a chunk of a step
//   step record: 0 K | 1 n_mul | 2 flags | 3 sum_size | 4 out_off | 5 otab_off | 6 stab_off | 7 wsmask | 8 evmask | 9 fast code
//                (fast code = K*8 + #leading const operands when the step is a plain sum-product whose const operands
//                 come first, else -1: selects a fully specialised instantiation of the chunk function)
//                | then per operand (n_ev, ev_pairs_off) ; otab[out_size][K] ; stab[sum_size][K] ; ev pairs (slot, stride)
//   segs[n_segs][4]: (ws offset, size, out offset, flags)        colmap[out_elems]: ws entry of each output column
#pragma once
#include <cstdint>
#include <vector>

#include "pgx_step.cuh"

namespace pgx {

constexpr int MW_HEADER = 8;
constexpr int SR_FIXED = 10;
constexpr int ITEM_WORDS = 3;
constexpr int FUSED_LANES = 32;

struct MicroInfo {
    bool ok = false;
    int n_levels = 0;
    int n_items = 0;
    int max_k = 0;
    int64_t words = 0;
    bool all_fast = false;  // every step has a fast code
};

// Host: expand the packed plan into the microprogram. Returns false when the plan is too large for tables.
inline bool build_micro(const int32_t* pool, std::vector<int32_t>& w, MicroInfo& info, int64_t max_words = 1 << 21,
                        int chunks_per_level = 16) {
    const int n_steps = pool[3], n_segs = pool[4], out_elems = pool[5];
    const int32_t* index = pool + pool[10];
    int n_levels = 0;
    int64_t total = 0;
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* r = pool + index[s];
        const int K = r[2];
        const int64_t out_size = ld_i64(r + 4), sum_size = ld_i64(r + 6);
        if (out_size > (1 << 20) || sum_size > (1 << 20)) return false;
        total += SR_FIXED + 2 * K + K * (out_size + sum_size) + 2 * 8 * K + ITEM_WORDS * out_size;
        if (r[10] + 1 > n_levels) n_levels = r[10] + 1;
        if (s > 0 && r[10] < pool[index[s - 1] + 10]) return false;  // steps must be sorted by level
    }
    total += MW_HEADER + n_levels + 1 + 4 * n_segs + out_elems;
    if (total > max_words) return false;
    // cost of every level: the chunks of a level should be of similar cost so the warps of a CTA finish together
    std::vector<double> level_cost(n_levels, 0.0);
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* r = pool + index[s];
        level_cost[r[10]] += (double)ld_i64(r + 4) * ((double)ld_i64(r + 6) * r[2] + 6.0);
    }
    w.clear();
    w.reserve((size_t)total);
    w.resize(MW_HEADER, 0);
    const int levels_off = (int)w.size();
    w.resize(w.size() + n_levels + 1, 0);
    std::vector<int32_t> items;  // appended after the step records
    std::vector<int32_t> level_first(n_levels + 1, 0);
    int cur_level = 0;
    int max_k = 0;
    bool all_fast = true;
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* r = pool + index[s];
        const int A = r[0], S = r[1], K = r[2], flags = r[3], level = r[10];
        const int64_t out_size = ld_i64(r + 4), sum_size = ld_i64(r + 6), out_off = ld_i64(r + 8);
        const int opw = OP_FIXED + A + S;
        const int32_t* odims = r + STEP_FIXED;
        const int32_t* sdims = odims + A;
        const int32_t* ops = sdims + S;
        while (cur_level < level) level_first[++cur_level] = (int32_t)(items.size() / ITEM_WORDS);
        if (K > max_k) max_k = K;
        int n_mul = K;
        while (n_mul > 0 && (ops[(n_mul - 1) * opw] & 0x100)) --n_mul;
        const int srec = (int)w.size();
        w.resize(w.size() + SR_FIXED + 2 * K, 0);
        w[srec + 0] = K;
        w[srec + 1] = n_mul;
        w[srec + 2] = flags;
        w[srec + 3] = (int32_t)sum_size;
        w[srec + 4] = (int32_t)out_off;
        int wsmask = 0, evmask = 0, n_const = 0;
        bool const_first = true;
        for (int k = 0; k < K; ++k) {
            const int32_t* op = ops + k * opw;
            if ((op[0] & 0xFF) == 1) {
                wsmask |= 1 << k;
            } else {
                if (n_const != k) const_first = false;
                ++n_const;
            }
            if (op[3] > 0) evmask |= 1 << k;
        }
        w[srec + 7] = wsmask;
        w[srec + 8] = evmask;
        w[srec + 9] = (flags == 0 && const_first && (K <= 4 || (K == 5 && n_const <= 2))) ? K * 8 + n_const : -1;
        if (w[srec + 9] < 0) all_fast = false;
        // otab[o][k]
        const int otab = (int)w.size();
        w[srec + 5] = otab;
        w.resize(w.size() + (size_t)K * out_size);
        std::vector<int32_t> digit(A > 0 ? A : 1, 0);
        for (int64_t o = 0; o < out_size; ++o) {
            for (int k = 0; k < K; ++k) {
                const int32_t* op = ops + k * opw;
                int64_t e = ld_i64(op + 1);
                for (int a = 0; a < A; ++a) e += (int64_t)digit[a] * op[OP_FIXED + a];
                w[otab + o * K + k] = (int32_t)e;
            }
            for (int a = A - 1; a >= 0; --a) {
                if (++digit[a] < odims[a]) break;
                digit[a] = 0;
            }
        }
        const int stab = (int)w.size();
        w[srec + 6] = stab;
        w.resize(w.size() + (size_t)K * sum_size);
        std::vector<int32_t> sd(S > 0 ? S : 1, 0);
        for (int64_t q = 0; q < sum_size; ++q) {
            for (int k = 0; k < K; ++k) {
                const int32_t* op = ops + k * opw;
                int64_t e = 0;
                for (int a = 0; a < S; ++a) e += (int64_t)sd[a] * op[OP_FIXED + A + a];
                w[stab + q * K + k] = (int32_t)e;
            }
            for (int a = S - 1; a >= 0; --a) {
                if (++sd[a] < sdims[a]) break;
                sd[a] = 0;
            }
        }
        for (int k = 0; k < K; ++k) {
            const int32_t* op = ops + k * opw;
            const int n_ev = op[3];
            w[srec + SR_FIXED + 2 * k] = n_ev;
            w[srec + SR_FIXED + 2 * k + 1] = (int32_t)w.size();
            for (int j = 0; j < n_ev; ++j) {
                w.push_back(r[op[4] + 2 * j]);
                w.push_back(r[op[4] + 2 * j + 1]);
            }
        }
        // cut the step into chunks of consecutive entries, sized by the level's cost
        const double step_cost = (double)out_size * ((double)sum_size * K + 6.0);
        int64_t n_chunks = (int64_t)(step_cost / (level_cost[level] / chunks_per_level) + 0.5);
        n_chunks = n_chunks < 1 ? 1 : (n_chunks > out_size ? out_size : n_chunks);
        const int64_t per = (out_size + n_chunks - 1) / n_chunks;
        for (int64_t o = 0; o < out_size; o += per) {
            items.push_back(srec);
            items.push_back((int32_t)o);
            items.push_back((int32_t)(o + per <= out_size ? per : out_size - o));
        }
    }
    while (cur_level < n_levels) level_first[++cur_level] = (int32_t)(items.size() / ITEM_WORDS);
    for (int l = 0; l <= n_levels; ++l) w[levels_off + l] = level_first[l];
    const int items_off = (int)w.size();
    w.insert(w.end(), items.begin(), items.end());
    const int segs_off = (int)w.size();
    const int32_t* segs = pool + pool[11];
    for (int g = 0; g < n_segs; ++g) {
        const int32_t* sg = segs + g * SEG_WORDS;
        w.push_back((int32_t)ld_i64(sg));
        w.push_back(sg[2]);
        w.push_back(sg[3]);
        w.push_back(sg[4]);
    }
    const int colmap_off = (int)w.size();
    w.resize(w.size() + out_elems, 0);
    for (int g = 0; g < n_segs; ++g) {
        const int32_t* sg = segs + g * SEG_WORDS;
        for (int i = 0; i < sg[2]; ++i) w[colmap_off + sg[3] + i] = (int32_t)ld_i64(sg) + i;
    }
    w[0] = n_levels;
    w[1] = (int32_t)(items.size() / ITEM_WORDS);
    w[2] = levels_off;
    w[3] = items_off;
    w[4] = n_segs;
    w[5] = segs_off;
    w[6] = colmap_off;
    w[7] = out_elems;
    info.ok = true;
    info.n_levels = n_levels;
    info.n_items = (int)(items.size() / ITEM_WORDS);
    info.max_k = max_k;
    info.words = (int64_t)w.size();
    info.all_fast = all_fast;
    return true;
}

#if defined(__CUDACC__)

#ifndef PGX_FUSED_FAST_MINB
#define PGX_FUSED_FAST_MINB 2  // 512-thread CTAs per SM the fast-only fused kernel is compiled for (64 registers)
#endif

// Fast path: plain sum-product step whose NC batch-invariant operands come first. Everything about the operand
// list is a compile-time constant, so the inner loop is: offset-table load, add, value load, multiply.
//   wsb   : work-table base for this lane (shared: ws_s + lane, global: ws_g + b)
//   pitch : elements between consecutive entries (32 in shared memory, ldb in global memory)
template <typename T, int K, int NC, bool SMEM>
__device__ __forceinline__ void micro_chunk_fast(const int32_t* __restrict__ mp, const int32_t* __restrict__ sr, int o0,
                                                 int n_o, const T* __restrict__ cst, T* wsb, int64_t pitch,
                                                 const int32_t* __restrict__ evs, int lane) {
    const int sum_size = __ldg(sr + 3);
    const int out_off = __ldg(sr + 4);
    const int32_t* otab = mp + __ldg(sr + 5);
    const int32_t* stab = mp + __ldg(sr + 6);
    const int evmask = __ldg(sr + 8);
    int32_t evo[NC > 0 ? NC : 1];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
        evo[k] = 0;
        if ((evmask >> k) & 1) {
            const int n_ev = __ldg(sr + SR_FIXED + 2 * k);
            const int32_t* pairs = mp + __ldg(sr + SR_FIXED + 2 * k + 1);
            for (int j = 0; j < n_ev; ++j) evo[k] += evs[__ldg(pairs + 2 * j) * FUSED_LANES + lane] * __ldg(pairs + 2 * j + 1);
        }
    }
    // OU output entries per iteration: a step never reads what it writes, but loads and stores go through the same
    // pointer, so the compiler may not hoist the loads of entry o + 1 above the store of entry o. Doing OU entries
    // before any store multiplies the independent loads in flight per warp (the kernel is latency bound: ~50 % issue)
    // and shares every offset-table read among them.
    constexpr int OU = (K <= 2) ? 4 : ((K == 3) ? 3 : 2);
    int o = o0;
    for (; o + OU <= o0 + n_o; o += OU) {
        const int32_t* ot = otab + o * K;
        int32_t bo[OU][K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int32_t e = (k < NC ? evo[k < NC ? k : 0] : 0);
#pragma unroll
            for (int u = 0; u < OU; ++u) bo[u][k] = __ldg(ot + u * K + k) + e;
        }
        T acc[OU];
#pragma unroll
        for (int u = 0; u < OU; ++u) acc[u] = (T)0;
        const int32_t* st = stab;
#pragma unroll 2
        for (int s = 0; s < sum_size; ++s, st += K) {
            T pr[OU];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int32_t so = __ldg(st + k);
#pragma unroll
                for (int u = 0; u < OU; ++u) {
                    const int32_t f = bo[u][k] + so;
                    const T v = (k < NC) ? __ldg(cst + f) : (SMEM ? wsb[f * FUSED_LANES] : wsb[(int64_t)f * pitch]);
                    pr[u] = (k == 0) ? v : pr[u] * v;
                }
            }
#pragma unroll
            for (int u = 0; u < OU; ++u) acc[u] += pr[u];
        }
#pragma unroll
        for (int u = 0; u < OU; ++u) {
            if (SMEM)
                wsb[(out_off + o + u) * FUSED_LANES] = acc[u];
            else
                wsb[(int64_t)(out_off + o + u) * pitch] = acc[u];
        }
    }
    for (; o < o0 + n_o; ++o) {
        const int32_t* ot = otab + o * K;
        int32_t boff[K];
#pragma unroll
        for (int k = 0; k < K; ++k) boff[k] = __ldg(ot + k) + (k < NC ? evo[k < NC ? k : 0] : 0);
        T acc = (T)0;
        const int32_t* st = stab;
#pragma unroll 2
        for (int s = 0; s < sum_size; ++s, st += K) {
            T prod;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int32_t off = boff[k] + __ldg(st + k);
                const T v = (k < NC) ? __ldg(cst + off) : (SMEM ? wsb[off * FUSED_LANES] : wsb[(int64_t)off * pitch]);
                prod = (k == 0) ? v : prod * v;
            }
            acc += prod;
        }
        if (SMEM)
            wsb[(out_off + o) * FUSED_LANES] = acc;
        else
            wsb[(int64_t)(out_off + o) * pitch] = acc;
    }
}

// General path (divisors, max-reduce, interleaved operand kinds, more than 5 operands).
//   K     : operand slots compiled in; GENERIC instantiations read the real count from the record
template <typename T, int K, bool SMEM, bool GENERIC = false>
__device__ __forceinline__ void micro_chunk(const int32_t* __restrict__ mp, const int32_t* __restrict__ sr, int o0, int n_o,
                                            const T* __restrict__ cst, T* wsb, int64_t pitch,
                                            const int32_t* __restrict__ evs, int lane) {
    const int kk = GENERIC ? __ldg(sr) : K;  // table row length
    const int n_mul = __ldg(sr + 1);
    const int flags = __ldg(sr + 2);
    const int sum_size = __ldg(sr + 3);
    const int out_off = __ldg(sr + 4);
    const int32_t* otab = mp + __ldg(sr + 5);
    const int32_t* stab = mp + __ldg(sr + 6);
    const int wsmask = __ldg(sr + 7);
    const int evmask = __ldg(sr + 8);
    int32_t evo[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        evo[k] = 0;
        if ((evmask >> k) & 1) {  // bits above the real operand count are never set
            const int n_ev = __ldg(sr + SR_FIXED + 2 * k);
            const int32_t* pairs = mp + __ldg(sr + SR_FIXED + 2 * k + 1);
            for (int j = 0; j < n_ev; ++j) evo[k] += evs[__ldg(pairs + 2 * j) * FUSED_LANES + lane] * __ldg(pairs + 2 * j + 1);
        }
    }
    const bool use_max = (flags & FLAG_MAX) != 0;
    for (int o = o0; o < o0 + n_o; ++o) {
        const int32_t* ot = otab + o * kk;
        int32_t boff[K];
#pragma unroll
        for (int k = 0; k < K; ++k) boff[k] = (!GENERIC || k < kk) ? __ldg(ot + k) + evo[k] : 0;
        T acc = use_max ? neg_inf<T>() : (T)0;
        const int32_t* st = stab;
        for (int s = 0; s < sum_size; ++s, st += kk) {
            T prod = (T)1;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < n_mul) {
                    const int32_t off = boff[k] + __ldg(st + k);
                    T v;
                    if ((wsmask >> k) & 1)
                        v = SMEM ? wsb[off * FUSED_LANES] : wsb[(int64_t)off * pitch];
                    else
                        v = __ldg(cst + off);
                    prod *= v;
                }
            }
            if (use_max)
                acc = prod > acc ? prod : acc;
            else
                acc += prod;
        }
        if (flags & FLAG_DIV) {
            T den = (T)1;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k >= n_mul && (!GENERIC || k < kk)) {
                    const int32_t off = boff[k];
                    den *= ((wsmask >> k) & 1) ? (SMEM ? wsb[off * FUSED_LANES] : wsb[(int64_t)off * pitch]) : __ldg(cst + off);
                }
            }
            const T r = acc / den;
            acc = (r != r) ? (T)0 : r;
        }
        if (SMEM)
            wsb[(out_off + o) * FUSED_LANES] = acc;
        else
            wsb[(int64_t)(out_off + o) * pitch] = acc;
    }
}

// blockDim.x = 32 * G. Shared memory: [ws_entries][32] T (SMEM only), then evidence [n_ev][32] int32.
// FAST_ONLY: every step of the plan has a fast code (plain sum-product, const operands first, <= 5 operands), so the
// general chunk function is not compiled in — fewer registers, two 512-thread CTAs per SM.
template <typename T, bool SMEM, bool FAST_ONLY>
__global__ void __launch_bounds__(512, FAST_ONLY ? PGX_FUSED_FAST_MINB : 1) k_plan_fused2(const int32_t* __restrict__ mp, const T* __restrict__ cst,
                                                     T* __restrict__ ws_g, const int32_t* __restrict__ ev,
                                                     const int32_t* __restrict__ ev_card, T* __restrict__ out, int n_ev,
                                                     int ws_entries, int64_t B, int64_t ldb,
                                                     const T* __restrict__ soft, const int32_t* __restrict__ inputs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* ws_s = reinterpret_cast<T*>(smem_raw);
    int32_t* evs = reinterpret_cast<int32_t*>(smem_raw + (SMEM ? (size_t)ws_entries * FUSED_LANES * sizeof(T) : 0));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int G = blockDim.x >> 5;
    const int64_t row0 = (int64_t)blockIdx.x * FUSED_LANES;
    const int64_t b = row0 + lane;
    // evidence rows of this CTA -> shared [slot][lane], clamped into range
    for (int i = threadIdx.x; i < FUSED_LANES * n_ev; i += blockDim.x) {
        const int bb = i / n_ev, sl = i - bb * n_ev;
        int32_t st = 0;
        if (row0 + bb < B) {
            st = ev[(row0 + bb) * n_ev + sl];
            const int32_t card = __ldg(ev_card + sl);
            st = st < 0 ? 0 : (st >= card ? card - 1 : st);
        }
        evs[sl * FUSED_LANES + bb] = st;
    }
    if (inputs != nullptr) {
        // batch-dependent input tables (soft evidence): the CTA's 32 rows of soft[B, in_elems] -> work entries
        const int n_in = __ldg(inputs), in_elems = __ldg(inputs + 1);
        for (int i = threadIdx.x; i < FUSED_LANES * in_elems; i += blockDim.x) {
            const int bb = i / in_elems, e = i - bb * in_elems;
            if (row0 + bb >= B) continue;
            const T v = soft[(row0 + bb) * in_elems + e];
            for (int j = 0; j < n_in; ++j) {
                const int32_t* r = inputs + 2 + 4 * j;
                const int in_off = __ldg(r + 3), size = __ldg(r + 2);
                if (e >= in_off && e < in_off + size) {
                    const int entry = __ldg(r) + (e - in_off);
                    if (SMEM) ws_s[entry * FUSED_LANES + bb] = v; else ws_g[(int64_t)entry * ldb + row0 + bb] = v;
                    break;
                }
            }
        }
    }
    __syncthreads();
    T* wsb = SMEM ? ws_s + lane : ws_g + (b < B ? b : 0);
    const int n_levels = __ldg(mp + 0);
    const int32_t* levels = mp + __ldg(mp + 2);
    const int32_t* items = mp + __ldg(mp + 3);
    for (int lv = 0; lv < n_levels; ++lv) {
        const int i1 = __ldg(levels + lv + 1);
        for (int i = __ldg(levels + lv) + warp; i < i1; i += G) {
            const int32_t* sr = mp + __ldg(items + ITEM_WORDS * i);
            const int o0 = __ldg(items + ITEM_WORDS * i + 1);
            const int n_o = __ldg(items + ITEM_WORDS * i + 2);
            if (SMEM || b < B) {
#define PGX_FAST(KK, NN) \
    case KK * 8 + NN: micro_chunk_fast<T, KK, NN, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                switch (__ldg(sr + 9)) {
                    PGX_FAST(1, 0) PGX_FAST(1, 1)
                    PGX_FAST(2, 0) PGX_FAST(2, 1) PGX_FAST(2, 2)
                    PGX_FAST(3, 0) PGX_FAST(3, 1) PGX_FAST(3, 2) PGX_FAST(3, 3)
                    PGX_FAST(4, 0) PGX_FAST(4, 1) PGX_FAST(4, 2) PGX_FAST(4, 3) PGX_FAST(4, 4)
                    PGX_FAST(5, 0) PGX_FAST(5, 1) PGX_FAST(5, 2)
                    default:
                        if (!FAST_ONLY) switch (__ldg(sr)) {
                            case 1: micro_chunk<T, 1, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            case 2: micro_chunk<T, 2, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            case 3: micro_chunk<T, 3, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            case 4: micro_chunk<T, 4, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            case 5: micro_chunk<T, 5, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            case 6: micro_chunk<T, 6, SMEM>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                            default: micro_chunk<T, MAX_OPS, SMEM, true>(mp, sr, o0, n_o, cst, wsb, ldb, evs, lane); break;
                        }
                }
#undef PGX_FAST
            }
        }
        __syncthreads();
    }
    // emit: normalise the segments in place (lane = evidence set), then write whole output rows
    const int n_segs = __ldg(mp + 4);
    const int32_t* segs = mp + __ldg(mp + 5);
    const int32_t* colmap = mp + __ldg(mp + 6);
    const int out_elems = __ldg(mp + 7);
    if (SMEM) {
        for (int g = warp; g < n_segs; g += G) {
            if (__ldg(segs + 4 * g + 3) & SEG_NORMALIZE) {
                T* src = wsb + __ldg(segs + 4 * g) * FUSED_LANES;
                const int n = __ldg(segs + 4 * g + 1);
                T sum = (T)0;
                for (int i = 0; i < n; ++i) sum += src[i * FUSED_LANES];
                for (int i = 0; i < n; ++i) src[i * FUSED_LANES] = src[i * FUSED_LANES] / sum;
            }
        }
        __syncthreads();
        // the CTA's 32 output rows are one contiguous range of out[]: coalesced stores
        const int64_t rows = (B - row0) < FUSED_LANES ? (B - row0) : FUSED_LANES;
        const int64_t total = rows * out_elems;
        T* dst = out + row0 * out_elems;
        for (int64_t idx = threadIdx.x; idx < total; idx += blockDim.x) {
            const int bb = (int)(idx / out_elems);
            const int j = (int)(idx - (int64_t)bb * out_elems);
            dst[idx] = ws_s[__ldg(colmap + j) * FUSED_LANES + bb];
        }
    } else if (b < B) {
        for (int g = warp; g < n_segs; g += G) {
            const T* src = wsb + (int64_t)__ldg(segs + 4 * g) * ldb;
            const int n = __ldg(segs + 4 * g + 1);
            T* dst = out + b * out_elems + __ldg(segs + 4 * g + 2);
            if (__ldg(segs + 4 * g + 3) & SEG_NORMALIZE) {
                T sum = (T)0;
                for (int i = 0; i < n; ++i) sum += src[(int64_t)i * ldb];
                for (int i = 0; i < n; ++i) dst[i] = src[(int64_t)i * ldb] / sum;
            } else {
                for (int i = 0; i < n; ++i) dst[i] = src[(int64_t)i * ldb];
            }
        }
    }
}

#endif  // __CUDACC__

}  // namespace pgx
