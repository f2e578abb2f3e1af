// pgx.cu — CUDA kernels (sm_100a) and the C-ABI of libpgx.so (include/pgx.h).
//
// Kernels
//   k_contract_step   K2: one plan step over (output entries x evidence sets); evidence gather (K1) and
//                     divide fused into operand addressing / epilogue. HBM-bound for large tables.
//   k_plan_fused      K5 for small models: the WHOLE plan (collect, distribute, marginals, normalise) in one
//                     launch; a CTA owns a row of 32 evidence sets (lane = evidence set), G warps split the
//                     output entries of each step, __syncthreads() orders dependent steps.
//   k_emit            K4: per-segment normalise (values / values.sum(), NaN when the sum is 0) + transpose of
//                     the [entry][b] work layout into the caller's out[b, :] rows.
//   k_evidence_gather K1 stand-alone (DiscreteFactor.reduce for a batch).
//   k_normalize       K4 stand-alone.
//
// Layout: work tables [entry][ldb], evidence set fastest => every warp access is 32 consecutive elements
// regardless of which variable is summed; index arithmetic is warp-uniform.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pgx.h"
#include "pgx_step.cuh"
#include "pgx_fused.cuh"
#include "pgx_spec.h"
#include "pgx_mm.h"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define PGX_CUDA(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(PGX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));          \
    } while (0)

using namespace pgx;

#ifndef PGX_TILE32_MINB
#define PGX_TILE32_MINB 8  // CTAs per SM for the 32-bit-addressed tile kernel: 32 registers, 64 warps/SM (measured 4..8: 8 is best even with 40 B of spills)
#endif
#ifndef PGX_TILE_MINB
#define PGX_TILE_MINB 5  // CTAs per SM the 64-bit tile kernel is compiled for (register cap 65536 / (256 * N)); measured 1..6
#endif

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
// K2. Thread = (run of `opt` consecutive output entries, evidence set); lanes of a warp = 32 consecutive evidence
// sets at the same entries, so every access is a 256-byte (fp64) row segment whatever the strides are.
template <typename T, int MAXK>
__global__ void __launch_bounds__(256) k_contract_step(const int32_t* __restrict__ pool, int rec_off, int rec_len,
                                                       int ev_card_off, const T* __restrict__ cst, T* __restrict__ ws,
                                                       const int32_t* __restrict__ ev, int n_ev, int64_t B, int64_t ldb,
                                                       int bt_log2, int opt) {
    extern __shared__ int32_t s_rec[];
    for (int i = threadIdx.x; i < rec_len; i += blockDim.x) s_rec[i] = pool[rec_off + i];
    __syncthreads();
    const uint32_t bt_mask = (1u << bt_log2) - 1u;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = (int64_t)blockIdx.y * (bt_mask + 1u) + (t & bt_mask);
    const uint64_t o0 = (t >> bt_log2) * (uint64_t)opt;
    const uint64_t out_size = (uint64_t)ld_i64(s_rec + 4);
    if (b >= B || o0 >= out_size) return;
    const uint64_t o1 = o0 + opt < out_size ? o0 + opt : out_size;
    T* out = ws + ld_i64(s_rec + 8) * ldb + b;
    if (MAXK <= 8) {
        contract_run<T, (MAXK <= 8 ? MAXK : 8)>(s_rec, cst, ws, ev + b * n_ev, pool + ev_card_off, ldb, b, (uint32_t)o0,
                                                (uint32_t)o1, out);
    } else {
        for (uint64_t o = o0; o < o1; ++o)
            out[(int64_t)o * ldb] =
                contract_elem<T, MAXK>(s_rec, cst, ws, ev + b * n_ev, pool + ev_card_off, ldb, b, (uint32_t)o);
    }
}

// One (step, tile shape) entry of a level-batched launch of k_contract_tile.
struct TileItem {
    int32_t rec_off, rec_len;
    int32_t TO;        // output entries per CTA tile
    int32_t btb;       // evidence-set tiles (of 32) per CTA
    int32_t b_blocks;  // CTAs along the batch
    int32_t blk_begin; // first linear CTA index of this step inside the launch
};

// Phase 1 of the tile-cooperative kernels: thread t decodes output entry tile0 + t ONCE for the whole CTA into
// s_otab[t][k] (entry offset of operand k at summed index 0) and the summed range into s_stab[q][k]. Lanes work on
// different entries, so no lane repeats another's mixed-radix arithmetic (ncu on munin: the per-thread decomposition
// cost ~150 issue slots per (entry, warp) when every lane decoded the same entry).
// `work_unit` != 0: offsets of work-table operands are stored already multiplied by it (elements per table entry = ldb),
// so the streaming loop adds them to a row base without a multiply (32-bit-addressed kernel only).
template <int MAXK>
__device__ __forceinline__ void tile_decode(const int32_t* __restrict__ s_rec, int32_t* __restrict__ s_otab,
                                            int32_t* __restrict__ s_stab, uint32_t tile0, int TO, uint32_t work_unit = 0) {
    const int A = s_rec[0], S = s_rec[1], K = s_rec[2];
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = s_rec + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;
    const uint32_t out_size = (uint32_t)s_rec[4];
    const int sum_size = s_rec[6];
    for (int t = threadIdx.x; t < TO; t += blockDim.x) {
        uint32_t rem = tile0 + t;
        int32_t off[MAXK];
#pragma unroll
        for (int k = 0; k < MAXK; ++k) off[k] = 0;
        if (rem < out_size) {
            for (int a = A - 1; a >= 0; --a) {
                const uint32_t d = (uint32_t)odims[a];
                const uint32_t q = rem / d;
                const int32_t digit = (int32_t)(rem - q * d);
                rem = q;
#pragma unroll
                for (int k = 0; k < MAXK; ++k)
                    if (k < K) off[k] += digit * ops[k * opw + OP_FIXED + a];
            }
        }
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K) s_otab[t * K + k] = (work_unit && (ops[k * opw] & 0xFF) == 1) ? (int32_t)((uint32_t)off[k] * work_unit) : off[k];
    }
    for (int qi = threadIdx.x; qi < sum_size; qi += blockDim.x) {
        uint32_t rem = (uint32_t)qi;
        int32_t off[MAXK];
#pragma unroll
        for (int k = 0; k < MAXK; ++k) off[k] = 0;
        for (int a = S - 1; a >= 0; --a) {
            const uint32_t d = (uint32_t)sdims[a];
            const uint32_t q = rem / d;
            const int32_t digit = (int32_t)(rem - q * d);
            rem = q;
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
                if (k < K) off[k] += digit * ops[k * opw + OP_FIXED + A + a];
        }
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K) s_stab[qi * K + k] = (work_unit && (ops[k * opw] & 0xFF) == 1) ? (int32_t)((uint32_t)off[k] * work_unit) : off[k];
    }
}

// which step of a level-batched launch does this CTA belong to? (items sorted by blk_begin)
template <typename Item>
__device__ __forceinline__ int find_item(const Item* __restrict__ items, int n_items) {
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].blk_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// K2, tile-cooperative form (the default for steps whose summed range fits the shared offset table).
// ncu on munin showed the per-thread mixed-radix decomposition costs ~150 issue slots per (entry, warp): SIMT lanes
// of a warp all decode the SAME entry. Here a CTA owns a tile of `TO` consecutive output entries and up to `btb`
// tiles of 32 evidence sets:
//   phase 1  thread t decodes entry tile0 + t ONCE for the whole CTA -> s_otab[t][k]; the summed range is decoded the
//            same way into s_stab[q][k]  (lanes work on different entries: no redundant index arithmetic);
//   phase 2  lanes = evidence sets again; a warp streams rows  sum_q prod_k operand_k[(otab + stab) * unit]  with
//            nothing but shared-memory offset reads, one IMAD.WIDE, the load and the multiply per operand.
// One launch covers ALL tile-eligible steps of a dependency level (they are independent): the linear CTA index is
// mapped to its step through `items` (binary search), so the thousands of tiny steps of a large junction tree cost
// one launch per level instead of one launch each.
template <typename T, int MAXK>
__global__ void __launch_bounds__(256, MAXK > 4 ? 1 : PGX_TILE_MINB) k_contract_tile(const int32_t* __restrict__ pool,
                                                       const TileItem* __restrict__ items, int n_items, int ev_card_off,
                                                       const T* __restrict__ cst, const T* __restrict__ ws_in,
                                                       T* __restrict__ ws_out, const int32_t* __restrict__ ev, int n_ev,
                                                       int64_t B, int64_t ldb, int bt_log2) {
    extern __shared__ int32_t s_mem[];
    const TileItem it = items[find_item(items, n_items)];
    const int local = (int)blockIdx.x - it.blk_begin;
    const int tile_x = local / it.b_blocks;
    const int b_block = local - tile_x * it.b_blocks;
    const int TO = it.TO, btb = it.btb, rec_len = it.rec_len;

    int32_t* s_rec = s_mem;
    for (int i = threadIdx.x; i < rec_len; i += blockDim.x) s_rec[i] = pool[it.rec_off + i];
    __syncthreads();
    const int A = s_rec[0], S = s_rec[1], K = s_rec[2], flags = s_rec[3];
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = s_rec + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;
    const uint32_t out_size = (uint32_t)s_rec[4];
    const int sum_size = s_rec[6];
    int32_t* s_otab = s_mem + ((rec_len + 3) & ~3);
    int32_t* s_stab = s_otab + TO * K;
    const uint32_t tile0 = (uint32_t)tile_x * (uint32_t)TO;
    tile_decode<MAXK>(s_rec, s_otab, s_stab, tile0, TO);
    __syncthreads();

    int n_mul = K;
    if (flags & FLAG_DIV)
        while (n_mul > 0 && (ops[(n_mul - 1) * opw] & 0x100)) --n_mul;
    const bool use_max = (flags & FLAG_MAX) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int bt = 1 << bt_log2;
    const int o_per_warp = 32 >> bt_log2;  // entries a warp covers at once when the batch is narrower than 32
    const int o_sub = lane >> bt_log2;
    const int64_t out_off = ld_i64(s_rec + 8);
    const int32_t* ev_card = pool + ev_card_off;
    const int og_step = n_warps * o_per_warp;
    for (int tb = 0; tb < btb; ++tb) {
        const int64_t b = ((int64_t)b_block * btb + tb) * bt + (lane & (bt - 1));
        if (b >= B) continue;  // no barriers below
        const T* base[MAXK];
        int64_t unit[MAXK];
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            base[k] = cst;
            unit[k] = 1;
            if (k < K) {
                const int32_t* op = ops + k * opw;
                int64_t e = ld_i64(op + 1);
                const int ne = op[3];
                if (ne > 0) {
                    const int32_t* pairs = s_rec + op[4];
                    for (int j = 0; j < ne; ++j) {
                        const int slot = pairs[2 * j];
                        int32_t st = ev[b * n_ev + slot];
                        const int32_t card = ev_card[slot];
                        st = st < 0 ? 0 : (st >= card ? card - 1 : st);
                        e += st * pairs[2 * j + 1];
                    }
                }
                if ((op[0] & 0xFF) == 1) {
                    unit[k] = ldb;
                    base[k] = ws_in + e * ldb + b;
                } else {
                    base[k] = cst + e;
                }
            }
        }
        T* out = ws_out + out_off * ldb + b;
        if (S == 0 && !(flags & FLAG_DIV)) {
            // pure product: two entries in flight per thread (the step never reads what it writes, so the loads of
            // the second entry may be issued before the store of the first)
            for (int og = warp * o_per_warp + o_sub; og < TO; og += 2 * og_step) {
                const uint32_t o0 = tile0 + og, o1 = o0 + og_step;
                if (o0 >= out_size) break;
                const bool two = (og + og_step < TO) && (o1 < out_size);
                const int32_t* ot0 = s_otab + og * K;
                const int32_t* ot1 = ot0 + (two ? og_step * K : 0);
                T v0[MAXK], v1[MAXK];
#pragma unroll
                for (int k = 0; k < MAXK; ++k) {
                    if (k < K) {
                        v0[k] = base[k][(int64_t)ot0[k] * unit[k]];
                        v1[k] = base[k][(int64_t)ot1[k] * unit[k]];
                    }
                }
                T p0 = (T)1, p1 = (T)1;
#pragma unroll
                for (int k = 0; k < MAXK; ++k) {
                    if (k < K) {
                        p0 *= v0[k];
                        p1 *= v1[k];
                    }
                }
                out[(int64_t)o0 * ldb] = p0;
                if (two) out[(int64_t)o1 * ldb] = p1;
            }
            continue;
        }
        for (int og = warp * o_per_warp + o_sub; og < TO; og += og_step) {
            const uint32_t o = tile0 + og;
            if (o >= out_size) break;
            const int32_t* ot = s_otab + og * K;
            const T* ptr[MAXK];
#pragma unroll
            for (int k = 0; k < MAXK; ++k) ptr[k] = (k < K) ? base[k] + (int64_t)ot[k] * unit[k] : base[k];
            T acc;
            if (S == 0) {
                T prod = (T)1;
#pragma unroll
                for (int k = 0; k < MAXK; ++k)
                    if (k < n_mul) prod *= *ptr[k];
                acc = prod;
            } else {
                acc = use_max ? neg_inf<T>() : (T)0;
                const int32_t* st = s_stab;
#pragma unroll 4
                for (int q = 0; q < sum_size; ++q, st += K) {
                    T prod = (T)1;
#pragma unroll
                    for (int k = 0; k < MAXK; ++k)
                        if (k < n_mul) prod *= ptr[k][(int64_t)st[k] * unit[k]];
                    if (use_max)
                        acc = prod > acc ? prod : acc;
                    else
                        acc += prod;
                }
            }
            if (flags & FLAG_DIV) {
                T den = (T)1;
#pragma unroll
                for (int k = 0; k < MAXK; ++k)
                    if (k >= n_mul && k < K) den *= *ptr[k];
                const T r = acc / den;
                acc = (r != r) ? (T)0 : r;
            }
            out[(int64_t)o * ldb] = acc;
        }
    }
}

// Phase 2 of k_contract_tile32 for one tile of evidence sets, specialised on the number of multiplicands NM: the
// offset tables already hold element offsets (tile_decode with work_unit = ldb), so one operand element costs an
// offset read, an add, the load and the multiply. (The first version ran ONE loop for every step shape: operand count
// as a run-time predicate on a MAXK-wide body, sum and max both computed and selected, one IMAD per load — ncu on
// munin: 38 issued instructions per (entry, summed index) of a two-operand step, 68 % issue-active on a kernel that
// should only wait for HBM; profiles/r02_tile32_munin_ncu_raw.txt.)
//   pm[k]  : element index of operand k's row for this lane at (entry offset 0, summed index 0)
//   d0, d1 : the same for up to two divisors (n_div of them), whose offsets are columns NM, NM + 1 of the entry table
template <typename T, int NM, bool MAX>
__device__ __forceinline__ void tile32_rows(const T* __restrict__ ws_in, T* __restrict__ ws_out,
                                            const int32_t* __restrict__ s_otab, const int32_t* __restrict__ s_stab, int K,
                                            int sum_size, const uint32_t (&pm)[NM], int n_div, uint32_t d0, uint32_t d1,
                                            uint32_t outb, uint32_t ldb, uint32_t tile0, uint32_t out_size, int TO, int og0,
                                            int og_step) {
    if (sum_size == 1 && n_div == 0) {
        // pure product: two entries in flight per thread (the step never reads what it writes, so the loads of the second
        // entry may be issued before the store of the first)
        for (int og = og0; og < TO; og += 2 * og_step) {
            const uint32_t o0 = tile0 + og, o1 = o0 + og_step;
            if (o0 >= out_size) break;
            const bool two = (og + og_step < TO) && (o1 < out_size);
            const int32_t* ot0 = s_otab + og * K;
            const int32_t* ot1 = ot0 + (two ? og_step * K : 0);
            T v0[NM], v1[NM];
#pragma unroll
            for (int k = 0; k < NM; ++k) {
                v0[k] = ws_in[pm[k] + (uint32_t)ot0[k]];
                v1[k] = ws_in[pm[k] + (uint32_t)ot1[k]];
            }
            T p0 = v0[0], p1 = v1[0];
#pragma unroll
            for (int k = 1; k < NM; ++k) {
                p0 *= v0[k];
                p1 *= v1[k];
            }
            ws_out[outb + o0 * ldb] = p0;
            if (two) ws_out[outb + o1 * ldb] = p1;
        }
        return;
    }
    for (int og = og0; og < TO; og += og_step) {
        const uint32_t o = tile0 + og;
        if (o >= out_size) break;
        const int32_t* ot = s_otab + og * K;
        uint32_t p[NM];
#pragma unroll
        for (int k = 0; k < NM; ++k) p[k] = pm[k] + (uint32_t)ot[k];
        T acc;
        if (sum_size == 1) {
            acc = ws_in[p[0]];
#pragma unroll
            for (int k = 1; k < NM; ++k) acc *= ws_in[p[k]];
        } else {
            acc = MAX ? neg_inf<T>() : (T)0;
            const int32_t* st = s_stab;
#pragma unroll 4
            for (int q = 0; q < sum_size; ++q, st += K) {
                T prod = ws_in[p[0] + (uint32_t)st[0]];
#pragma unroll
                for (int k = 1; k < NM; ++k) prod *= ws_in[p[k] + (uint32_t)st[k]];
                if (MAX)
                    acc = prod > acc ? prod : acc;
                else
                    acc += prod;
            }
        }
        if (n_div > 0) {
            T den = ws_in[d0 + (uint32_t)ot[NM]];
            if (n_div > 1) den *= ws_in[d1 + (uint32_t)ot[NM + 1]];
            const T r = acc / den;
            acc = (r != r) ? (T)0 : r;  // 0/0 -> 0 ; x/0 stays inf (DiscreteFactor.py:859-863)
        }
        ws_out[outb + o * ldb] = acc;
    }
}

// K2, tile-cooperative form with 32-bit addressing (the default). Same two phases as k_contract_tile, but every
// operand lives in ONE address space: pgx_run_batch copies the batch-invariant tables to the head of the workspace,
// so an operand element is `wsb[u32 index]` whatever its kind. Per operand the thread keeps one 32-bit row base, which
// brings the kernel to 32 registers — these steps wait for HBM, so resident warps are what buys throughput. The host
// sends a step here only if it has at most two divisors (else the generic kernel takes it).
template <typename T, int MAXK>
__global__ void __launch_bounds__(256, MAXK <= 4 ? PGX_TILE32_MINB : (MAXK <= 6 ? 5 : 3)) k_contract_tile32(const int32_t* __restrict__ pool,
                                                                            const TileItem* __restrict__ items, int n_items,
                                                                            int ev_card_off, const T* __restrict__ ws_in,
                                                                            T* __restrict__ ws_out, uint32_t ws_off0,
                                                                            const int32_t* __restrict__ ev, int n_ev,
                                                                            int64_t B, uint32_t ldb, int bt_log2) {
    extern __shared__ int32_t s_mem[];
    const TileItem it = items[find_item(items, n_items)];
    const int local = (int)blockIdx.x - it.blk_begin;
    const int tile_x = local / it.b_blocks;
    const int b_block = local - tile_x * it.b_blocks;
    const int TO = it.TO, btb = it.btb, rec_len = it.rec_len;

    int32_t* s_rec = s_mem;
    for (int i = threadIdx.x; i < rec_len; i += blockDim.x) s_rec[i] = pool[it.rec_off + i];
    __syncthreads();
    const int A = s_rec[0], S = s_rec[1], K = s_rec[2], flags = s_rec[3];
    const int opw = OP_FIXED + A + S;
    const int32_t* ops = s_rec + STEP_FIXED + A + S;
    const uint32_t out_size = (uint32_t)s_rec[4];
    const int sum_size = s_rec[6];
    int32_t* s_otab = s_mem + ((rec_len + 3) & ~3);
    int32_t* s_stab = s_otab + TO * K;
    const uint32_t tile0 = (uint32_t)tile_x * (uint32_t)TO;
    tile_decode<MAXK>(s_rec, s_otab, s_stab, tile0, TO, ldb);
    __syncthreads();

    int n_mul = K;
    if (flags & FLAG_DIV)
        while (n_mul > 0 && (ops[(n_mul - 1) * opw] & 0x100)) --n_mul;
    const int n_div = K - n_mul;
    const bool use_max = (flags & FLAG_MAX) != 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const int bt = 1 << bt_log2;
    const int o_per_warp = 32 >> bt_log2;
    const int o_sub = lane >> bt_log2;
    const uint32_t out_base = ws_off0 + (uint32_t)s_rec[8] * ldb;  // out work offset fits 32 bits here (host checked)
    const int32_t* ev_card = pool + ev_card_off;
    const int og_step = n_warps * o_per_warp;
    const int og0 = warp * o_per_warp + o_sub;
    for (int tb = 0; tb < btb; ++tb) {
        const int64_t b = ((int64_t)b_block * btb + tb) * bt + (lane & (bt - 1));
        if (b >= B) continue;  // no barriers below
        uint32_t rowb[MAXK];
        uint32_t d0 = 0, d1 = 0;
#pragma unroll
        for (int k = 0; k < MAXK; ++k) {
            rowb[k] = 0;
            if (k < K) {
                const int32_t* op = ops + k * opw;
                uint32_t e = (uint32_t)op[1];
                const int ne = op[3];
                if (ne > 0) {
                    const int32_t* pairs = s_rec + op[4];
                    for (int j = 0; j < ne; ++j) {
                        const int slot = pairs[2 * j];
                        int32_t st = ev[b * n_ev + slot];
                        const int32_t card = ev_card[slot];
                        st = st < 0 ? 0 : (st >= card ? card - 1 : st);
                        e += (uint32_t)(st * pairs[2 * j + 1]);
                    }
                }
                rowb[k] = (op[0] & 0xFF) == 1 ? ws_off0 + e * ldb + (uint32_t)b : e;
                if (k == n_mul) d0 = rowb[k];
                if (k == n_mul + 1) d1 = rowb[k];
            }
        }
        const uint32_t outb = out_base + (uint32_t)b;
#define PGX_T32_CASE(NM)                                                                                                  \
    case NM: {                                                                                                            \
        if (NM <= MAXK) {                                                                                                 \
            uint32_t pm[NM];                                                                                              \
            _Pragma("unroll") for (int k = 0; k < NM; ++k) pm[k] = rowb[k < MAXK ? k : 0];                                \
            if (use_max)                                                                                                  \
                tile32_rows<T, NM, true>(ws_in, ws_out, s_otab, s_stab, K, sum_size, pm, n_div, d0, d1, outb, ldb, tile0, \
                                         out_size, TO, og0, og_step);                                                     \
            else                                                                                                          \
                tile32_rows<T, NM, false>(ws_in, ws_out, s_otab, s_stab, K, sum_size, pm, n_div, d0, d1, outb, ldb, tile0, \
                                          out_size, TO, og0, og_step);                                                    \
        }                                                                                                                 \
    } break;
        switch (n_mul) {
            PGX_T32_CASE(1)
            PGX_T32_CASE(2)
            PGX_T32_CASE(3)
            PGX_T32_CASE(4)
            PGX_T32_CASE(5)
            PGX_T32_CASE(6)
            PGX_T32_CASE(7)
            PGX_T32_CASE(8)
            default: break;
        }
#undef PGX_T32_CASE
    }
}

template <typename T>
__device__ __forceinline__ void emit_segment(const int32_t* __restrict__ seg, const T* __restrict__ ws,
                                             T* __restrict__ out, int64_t out_elems, int64_t ldb, int64_t b) {
    const int64_t off = ld_i64(seg);
    const int n = seg[2];
    const T* src = ws + off * ldb + b;
    T* dst = out + b * out_elems + seg[3];
    if (seg[4] & SEG_NORMALIZE) {
        T sum = (T)0;
        for (int i = 0; i < n; ++i) sum += src[(int64_t)i * ldb];
        for (int i = 0; i < n; ++i) dst[i] = src[(int64_t)i * ldb] / sum;
    } else {
        for (int i = 0; i < n; ++i) dst[i] = src[(int64_t)i * ldb];
    }
}

template <typename T>
__global__ void __launch_bounds__(128) k_emit(const int32_t* __restrict__ pool, int segs_off, const T* __restrict__ ws,
                                              T* __restrict__ out, int64_t out_elems, int64_t B, int64_t ldb) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    emit_segment<T>(pool + segs_off + blockIdx.y * SEG_WORDS, ws, out, out_elems, ldb, b);
}

// Whole plan in one launch. blockDim = (32, G): lane = evidence set within the row, G warps share the row.
template <typename T, int MAXK>
__global__ void __launch_bounds__(512) k_plan_fused(const int32_t* __restrict__ pool, const T* __restrict__ cst,
                                                     T* __restrict__ ws, const int32_t* __restrict__ ev,
                                                     T* __restrict__ out, int64_t B, int64_t ldb) {
    const int n_ev = pool[2];
    const int n_steps = pool[3];
    const int n_segs = pool[4];
    const int64_t out_elems = pool[5];
    const int32_t* index = pool + pool[10];
    const int32_t* ev_card = pool + pool[14];
    const int G = blockDim.y;
    const int w = threadIdx.y;
    const int64_t b = (int64_t)blockIdx.x * 32 + threadIdx.x;
    const bool live = b < B;
    const int32_t* ev_row = ev + (live ? b : 0) * n_ev;
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* rec = pool + index[s];
        if (live) {
            const uint32_t out_size = (uint32_t)rec[4];
            const int64_t out_off = ld_i64(rec + 8);
            for (uint32_t o = w; o < out_size; o += G) {
                const T v = contract_elem_upto<T, MAXK>(rec, cst, ws, ev_row, ev_card, ldb, b, o);
                ws[(out_off + o) * ldb + b] = v;
            }
        }
        if (G > 1) __syncthreads();
    }
    if (live) {
        const int32_t* segs = pool + pool[11];
        for (int g = w; g < n_segs; g += G) emit_segment<T>(segs + g * SEG_WORDS, ws, out, out_elems, ldb, b);
    }
}

// dst[e, b] = table[base(e) + sum_j clamp(ev[b, slot_j]) * stride_j]
template <typename T>
__global__ void __launch_bounds__(256) k_evidence_gather(const T* __restrict__ table, int n_free, const int32_t* dims,
                                                         const int32_t* strides, int n_ev, const int32_t* slots,
                                                         const int32_t* ev_strides, const int32_t* ev_cards,
                                                         const int32_t* __restrict__ ev, int ev_row_len,
                                                         T* __restrict__ dst, int64_t n_entries, int64_t B, int64_t ldb,
                                                         int bt_log2) {
    const uint32_t bt_mask = (1u << bt_log2) - 1u;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t b = (int64_t)blockIdx.y * (bt_mask + 1u) + (t & bt_mask);
    const uint64_t e = t >> bt_log2;
    if (b >= B || e >= (uint64_t)n_entries) return;
    int64_t off = 0;
    uint32_t rem = (uint32_t)e;
    for (int a = n_free - 1; a >= 0; --a) {
        const uint32_t d = (uint32_t)dims[a];
        const uint32_t q = rem / d;
        off += (int64_t)(rem - q * d) * strides[a];
        rem = q;
    }
    for (int j = 0; j < n_ev; ++j) {
        int32_t st = ev[b * ev_row_len + slots[j]];
        st = st < 0 ? 0 : (st >= ev_cards[j] ? ev_cards[j] - 1 : st);
        off += (int64_t)st * ev_strides[j];
    }
    dst[(int64_t)e * ldb + b] = table[off];
}

// Batch-dependent input tables (soft evidence): ws[(off_j + i), b] = soft[b, in_off_j + i]. One launch for all inputs of
// the plan: thread = (evidence set, element of the input row), reads coalesced along the row.
template <typename T>
__global__ void __launch_bounds__(256) k_scatter_inputs(const int32_t* __restrict__ inputs, const T* __restrict__ soft,
                                                        T* __restrict__ ws, int64_t B, int64_t ldb) {
    const int n_in = inputs[0], in_elems = inputs[1];
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * in_elems) return;
    const int64_t b = idx / in_elems;
    const int e = (int)(idx - b * in_elems);
    for (int j = 0; j < n_in; ++j) {
        const int32_t* r = inputs + 2 + 4 * j;
        if (e >= r[3] && e < r[3] + r[2]) {
            ws[(ld_i64(r) + (e - r[3])) * ldb + b] = soft[idx];
            return;
        }
    }
}

// Max-product traceback (most probable explanation). The plan left one "upward belief" beta_i per clique in the
// workspace (compile_jt_mpe_plan); walking the cliques root first, the variables of clique i that no ancestor fixed take
// the argmax of beta_i restricted to the ones already fixed (first maximum in C-order, like numpy.argmax over the joint,
// pgmpy/inference/ExactInference.py:609-612). Thread = evidence set: the restriction differs per evidence set, so the
// reads are gathers; cliques are small next to the message passing that built them.
//   trace: n_cliques | n_columns | per clique: work offset lo, hi | n_axes | n_axes x (column, card, stride, is_new)
template <typename T>
__global__ void __launch_bounds__(128) k_mpe_traceback(const int32_t* __restrict__ trace, const T* __restrict__ ws,
                                                       int32_t* __restrict__ assign, int64_t B, int64_t ldb) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int n_cliques = trace[0], n_cols = trace[1];
    int32_t* mine = assign + b * n_cols;
    const int32_t* r = trace + 2;
    for (int c = 0; c < n_cliques; ++c) {
        const int64_t off = ld_i64(r);
        const int n_ax = r[2];
        const int32_t* ax = r + 3;
        int64_t base = off;
        int64_t n_new = 1;
        for (int a = 0; a < n_ax; ++a) {
            if (ax[4 * a + 3]) n_new *= ax[4 * a + 1];
            else base += (int64_t)mine[ax[4 * a]] * ax[4 * a + 2];
        }
        // enumerate the new axes in C-order (last new axis fastest) with an odometer over their digits
        int digit[MAX_AXES];
        for (int a = 0; a < n_ax; ++a) digit[a] = 0;
        T best = neg_inf<T>();
        int64_t arg = 0, rel = 0;
        for (int64_t k = 0; k < n_new; ++k) {
            const T v = ws[(base + rel) * ldb + b];
            if (v > best || (k == 0)) {
                best = v;
                arg = k;
            }
            for (int a = n_ax - 1; a >= 0; --a) {
                if (!ax[4 * a + 3]) continue;
                rel += ax[4 * a + 2];
                if (++digit[a] < ax[4 * a + 1]) break;
                rel -= (int64_t)digit[a] * ax[4 * a + 2];
                digit[a] = 0;
            }
        }
        for (int a = n_ax - 1; a >= 0; --a) {
            if (!ax[4 * a + 3]) continue;
            const int d = ax[4 * a + 1];
            mine[ax[4 * a]] = (int32_t)(arg % d);
            arg /= d;
        }
        r += 3 + 4 * n_ax;
    }
}

// first index of the row maximum (numpy.argmax semantics: first occurrence; NaN propagates like numpy: a NaN wins)
template <typename T>
__global__ void __launch_bounds__(128) k_argmax_rows(const T* __restrict__ src, int64_t n, int64_t B, int32_t* __restrict__ out) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const T* row = src + b * n;
    T best = row[0];
    int32_t arg = 0;
    for (int64_t i = 1; i < n; ++i) {
        const T v = row[i];
        if (best == best && (v > best || v != v)) {
            best = v;
            arg = (int32_t)i;
        }
    }
    out[b] = arg;
}

template <typename T>
__global__ void __launch_bounds__(128) k_normalize(const T* __restrict__ src, int64_t n, int64_t ldb, T* __restrict__ out,
                                                   int64_t out_row_len, int64_t B) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    T sum = (T)0;
    for (int64_t i = 0; i < n; ++i) sum += src[i * ldb + b];
    for (int64_t i = 0; i < n; ++i) out[b * out_row_len + i] = src[i * ldb + b] / sum;
}

struct StepInfo {
    int rec_off;
    int rec_len;
    int64_t out_size;
    int64_t sum_size;
    int n_ops;
    int level;
};

int ilog2_floor(int64_t x) {
    int l = 0;
    while ((1LL << (l + 1)) <= x) ++l;
    return l;
}

}  // namespace

struct LaunchGroup {
    int generic_step = -1;  // >= 0: one launch of the generic kernel for this step
    int first_item = 0, n_items = 0, n_blocks = 0, max_k = 0;
    size_t smem = 0;
    bool mm = false;     // every step of the group goes to k_contract_mm (pgx_mm.cu: pipelined matrix-product tiles)
    bool tc = false;     // every step of the group goes to k_contract_tc32 (pgx_tc32.cu: tcgen05 TF32x3, fp32 mode)
    int level = 0;       // dependency level of its steps: launches of one level are independent of each other
    std::vector<int> step_ids;  // plan steps served by this launch (tracing: pgx_profile_launches)
};

struct StepSchedule {
    int64_t B = 0;
    int step_kernel = 0, dtype_size = 0;
    std::vector<LaunchGroup> groups;
    TileItem* d_items = nullptr;
    pgx::MMItem* d_mm_items = nullptr;
    int32_t* d_mm_tabs = nullptr;
    void release() {
        if (d_items) cudaFree(d_items);
        if (d_mm_items) cudaFree(d_mm_items);
        if (d_mm_tabs) cudaFree(d_mm_tabs);
        d_items = nullptr;
        d_mm_items = nullptr;
        d_mm_tabs = nullptr;
    }
};

struct GraphEntry {
    int64_t B;
    const void* ev;
    const void* soft;
    void* out;
    void* ws;
    int step_kernel;
    int dtype_size;
    int n_launches;
    cudaGraphExec_t exec;
};

struct pgx_plan {
    int dtype = PGX_F64;
    int device = 0;
    int n_ev = 0, n_steps = 0, n_segs = 0, out_elems = 0;
    int64_t ws_entries = 0, table_entries = 0;
    int step_index_off = 0, segs_off = 0, ev_card_off = 0, max_ops = 0;
    int inputs_off = 0, n_inputs = 0, in_elems = 0;  // batch-dependent input tables (soft evidence)
    std::vector<int32_t> pool;
    std::vector<StepInfo> steps;
    int32_t* d_pool = nullptr;
    const void* blob = nullptr;
    int64_t max_joint = 0;  // max over steps of out_size * sum_size
    // table-driven fused kernel (pgx_fused.cuh)
    pgx::MicroInfo micro;
    int32_t* d_micro = nullptr;
    // plan-specialised kernel (pgx_spec.cu), built on request by pgx_plan_specialize
    pgx::SpecKernel* spec = nullptr;
    // max-product traceback descriptor (pgx_plan_set_trace)
    int32_t* d_trace = nullptr;
    int trace_cols = 0;
    // options
    int mode = PGX_MODE_AUTO;
    int fused_warps = 0;  // 0 = auto
    int use_graph = 1;    // stepwise: replay the step sequence as a CUDA graph
    std::vector<GraphEntry> graphs;
    std::vector<GraphEntry> graph_candidates;  // argument tuples of recent graph-cache misses (exec unused)
    std::vector<StepSchedule> schedules;
    cudaStream_t cap_stream = nullptr;
    // launches of one dependency level are independent: the second one runs on an auxiliary stream (fork / join by events,
    // captured as parallel branches of the CUDA graph), so the tail of one overlaps the body of the other
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int level_streams = 1;
    int last_graph = 0;
    int stage = 1;        // matrix-product-shaped two-operand steps: 1 k_contract_mm (default), 0 streaming kernel only
                          // (PGX_OPT_STAGE)
    int mma = 1;          // fp64 tensor cores (DMMA) for steps whose P operand is batch invariant (PGX_OPT_MMA)
    int tc32 = 1;         // fp32 mode: tcgen05 TF32x3 GEMM for steps whose P operand is batch invariant (PGX_OPT_TC32)
    int batch_levels = 1; // share one launch among the tile-eligible steps of a dependency level
    cudaEvent_t* prof_events = nullptr;  // set only inside pgx_profile_steps
    int step_kernel = 0;  // 0 = auto (tile-cooperative where possible), 1 = generic per-thread kernel only
    int fused_kernel = 0; // 0 = auto, 1 = generic (v1), 2 = table-driven/shared workspace, 3 = table-driven/global workspace
    int last_variant = 0;
    int last_sched = -1;  // index into `schedules` of the most recent stepwise run
    // info
    int64_t last_launches = 0;
    int last_mode = 0;
};

extern "C" {

const char* pgx_last_error(void) { return g_err.c_str(); }
int32_t pgx_abi_version(void) { return PGX_ABI_VERSION; }

int64_t pgx_batch_ld(int64_t B) {
    if (B <= 0) return 1;
    if (B >= 32) return (B + 31) / 32 * 32;
    int64_t p = 1;
    while (p < B) p <<= 1;
    return p;
}

int pgx_plan_create(const pgx_plan_desc* desc, pgx_plan** out) {
    if (!desc || !out) return fail(PGX_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->abi_version != PGX_ABI_VERSION) return fail(PGX_ERR_INVALID, "ABI version mismatch");
    if (desc->dtype != PGX_F64 && desc->dtype != PGX_F32) return fail(PGX_ERR_INVALID, "unknown dtype");
    if (!desc->pool || desc->pool_words < HEADER_WORDS) return fail(PGX_ERR_INVALID, "pool too short");
    const int32_t* p = desc->pool;
    const int64_t W = desc->pool_words;
    if ((uint32_t)p[0] != MAGIC || p[1] != 1) return fail(PGX_ERR_INVALID, "bad magic/version in plan pool");
    pgx_plan* pl = new pgx_plan();
    pl->dtype = desc->dtype;
    pl->n_ev = p[2];
    pl->n_steps = p[3];
    pl->n_segs = p[4];
    pl->out_elems = p[5];
    pl->ws_entries = ld_i64(p + 6);
    pl->table_entries = desc->table_entries;
    pl->step_index_off = p[10];
    pl->segs_off = p[11];
    pl->max_ops = p[12];
    pl->ev_card_off = p[14];
    pl->blob = desc->table_blob;
    auto bad = [&](int code, const std::string& m) {
        delete pl;
        return fail(code, m);
    };
    if (pl->n_ev < 0 || pl->n_steps < 0 || pl->n_segs < 0 || pl->out_elems < 0 || pl->ws_entries < 1)
        return bad(PGX_ERR_INVALID, "negative count in plan header");
    if (ld_i64(p + 8) > desc->table_entries) return bad(PGX_ERR_BOUNDS, "table blob smaller than the plan expects");
    if (pl->ev_card_off < HEADER_WORDS || (int64_t)pl->ev_card_off + pl->n_ev > W)
        return bad(PGX_ERR_INVALID, "evidence cardinalities out of pool");
    if (pl->step_index_off < HEADER_WORDS || (int64_t)pl->step_index_off + pl->n_steps > W)
        return bad(PGX_ERR_INVALID, "step index out of pool");
    if (pl->segs_off < HEADER_WORDS || (int64_t)pl->segs_off + (int64_t)pl->n_segs * SEG_WORDS > W)
        return bad(PGX_ERR_INVALID, "segments out of pool");
    if (!desc->table_blob && pl->n_steps > 0) return bad(PGX_ERR_INVALID, "null table blob");
    const int32_t* ev_card = p + pl->ev_card_off;
    for (int j = 0; j < pl->n_ev; ++j)
        if (ev_card[j] < 1) return bad(PGX_ERR_INVALID, "evidence cardinality < 1");
    for (int s = 0; s < pl->n_steps; ++s) {
        const int64_t off = p[pl->step_index_off + s];
        if (off < HEADER_WORDS || off + STEP_FIXED > W) return bad(PGX_ERR_INVALID, "step record out of pool");
        const int32_t* r = p + off;
        const int A = r[0], S = r[1], K = r[2];
        if (A < 0 || S < 0 || A + S > MAX_AXES) return bad(PGX_ERR_UNSUPPORTED, "too many axes in a step");
        if (K < 1 || K > MAX_OPS) return bad(PGX_ERR_UNSUPPORTED, "operand count out of range");
        const int opw = OP_FIXED + A + S;
        int64_t len = STEP_FIXED + A + S + (int64_t)K * opw;
        if (off + len > W) return bad(PGX_ERR_INVALID, "step operands out of pool");
        const int64_t out_size = ld_i64(r + 4), sum_size = ld_i64(r + 6), out_off = ld_i64(r + 8);
        int64_t po = 1, ps = 1;
        for (int a = 0; a < A; ++a) {
            if (r[STEP_FIXED + a] < 1) return bad(PGX_ERR_INVALID, "axis extent < 1");
            po *= r[STEP_FIXED + a];
        }
        for (int a = 0; a < S; ++a) {
            if (r[STEP_FIXED + A + a] < 1) return bad(PGX_ERR_INVALID, "axis extent < 1");
            ps *= r[STEP_FIXED + A + a];
        }
        if (po != out_size || ps != sum_size) return bad(PGX_ERR_INVALID, "axis extents disagree with sizes");
        if (out_size < 1 || out_size >= (1LL << 31) || sum_size >= (1LL << 31))
            return bad(PGX_ERR_UNSUPPORTED, "step index space too large");
        if (out_off < 0 || out_off + out_size > pl->ws_entries) return bad(PGX_ERR_BOUNDS, "step output outside workspace");
        bool seen_div = false;
        for (int k = 0; k < K; ++k) {
            const int32_t* op = r + STEP_FIXED + A + S + k * opw;
            const int kind = op[0] & 0xFF;
            const bool div = (op[0] & 0x100) != 0;
            if (kind != 0 && kind != 1) return bad(PGX_ERR_INVALID, "unknown operand kind");
            if (seen_div && !div) return bad(PGX_ERR_INVALID, "divisors must be the trailing operands");
            seen_div = seen_div || div;
            if (div && !(r[3] & FLAG_DIV)) return bad(PGX_ERR_INVALID, "divisor without FLAG_DIV");
            const int64_t base = ld_i64(op + 1);
            int64_t hi = base;
            for (int a = 0; a < A; ++a) {
                if (op[OP_FIXED + a] < 0) return bad(PGX_ERR_INVALID, "negative stride");
                hi += (int64_t)(r[STEP_FIXED + a] - 1) * op[OP_FIXED + a];
            }
            for (int a = 0; a < S; ++a) {
                if (op[OP_FIXED + A + a] < 0) return bad(PGX_ERR_INVALID, "negative stride");
                if (div && op[OP_FIXED + A + a] != 0) return bad(PGX_ERR_INVALID, "divisor depends on a summed axis");
                hi += (int64_t)(r[STEP_FIXED + A + a] - 1) * op[OP_FIXED + A + a];
            }
            const int n_ev = op[3];
            const int64_t evo = op[4];
            if (n_ev < 0 || (n_ev > 0 && (evo < len || off + evo + 2LL * n_ev > W)))
                return bad(PGX_ERR_INVALID, "evidence pairs out of pool");
            for (int j = 0; j < n_ev; ++j) {
                const int slot = r[evo + 2 * j], stride = r[evo + 2 * j + 1];
                if (slot < 0 || slot >= pl->n_ev || stride < 0) return bad(PGX_ERR_INVALID, "bad evidence pair");
                hi += (int64_t)(ev_card[slot] - 1) * stride;
            }
            if (n_ev > 0) len = std::max<int64_t>(len, evo + 2LL * n_ev);
            const int64_t limit = kind == 1 ? pl->ws_entries : pl->table_entries;
            if (base < 0 || hi >= limit) return bad(PGX_ERR_BOUNDS, "operand range outside its table space");
            if (kind == 1) {
                // a step may not read what it writes
                const int64_t lo = base;
                if (!(hi < out_off || lo >= out_off + out_size))
                    return bad(PGX_ERR_BOUNDS, "step output overlaps one of its operands");
            }
        }
        pl->steps.push_back(StepInfo{(int)off, (int)len, out_size, sum_size, K, r[10]});
        pl->max_joint = std::max(pl->max_joint, out_size * std::max<int64_t>(1, sum_size));
    }
    int64_t out_total = 0;
    for (int g = 0; g < pl->n_segs; ++g) {
        const int32_t* sg = p + pl->segs_off + g * SEG_WORDS;
        const int64_t off = ld_i64(sg);
        if (off < 0 || sg[2] < 0 || off + sg[2] > pl->ws_entries) return bad(PGX_ERR_BOUNDS, "segment outside workspace");
        if (sg[3] < 0 || sg[3] + sg[2] > pl->out_elems) return bad(PGX_ERR_BOUNDS, "segment outside output row");
        out_total += sg[2];
    }
    (void)out_total;
    pl->inputs_off = p[15];
    if (pl->inputs_off != 0) {
        if (pl->inputs_off < HEADER_WORDS || (int64_t)pl->inputs_off + 2 > W) return bad(PGX_ERR_INVALID, "inputs block out of pool");
        pl->n_inputs = p[pl->inputs_off];
        pl->in_elems = p[pl->inputs_off + 1];
        if (pl->n_inputs < 1 || pl->in_elems < 1 || (int64_t)pl->inputs_off + 2 + 4LL * pl->n_inputs > W)
            return bad(PGX_ERR_INVALID, "inputs block out of pool");
        for (int j = 0; j < pl->n_inputs; ++j) {
            const int32_t* r = p + pl->inputs_off + 2 + 4 * j;
            const int64_t off = ld_i64(r);
            if (off < 0 || r[2] < 1 || off + r[2] > pl->ws_entries) return bad(PGX_ERR_BOUNDS, "input table outside workspace");
            if (r[3] < 0 || r[3] + r[2] > pl->in_elems) return bad(PGX_ERR_BOUNDS, "input table outside the input row");
        }
    }
    pl->pool.assign(p, p + W);
    std::vector<int32_t> micro_words;
    if (pl->n_steps > 0 && pl->max_joint <= (1 << 16)) {
        int cpl = 16;  // ~ one chunk per warp of a 16-warp CTA: decode overhead beats perfect balance (measured 12..96)
        if (const char* e = std::getenv("PGX_CHUNKS_PER_LEVEL")) cpl = std::max(1, std::atoi(e));  // tuning knob
        build_micro(p, micro_words, pl->micro, 1 << 21, cpl);
    }
    cudaError_t e = cudaGetDevice(&pl->device);
    if (e != cudaSuccess) return bad(PGX_ERR_CUDA, std::string("cudaGetDevice: ") + cudaGetErrorString(e));
    e = cudaMalloc((void**)&pl->d_pool, (size_t)W * sizeof(int32_t));
    if (e != cudaSuccess) return bad(PGX_ERR_CUDA, std::string("cudaMalloc(pool): ") + cudaGetErrorString(e));
    e = cudaMemcpy(pl->d_pool, p, (size_t)W * sizeof(int32_t), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(pl->d_pool);
        return bad(PGX_ERR_CUDA, std::string("cudaMemcpy(pool): ") + cudaGetErrorString(e));
    }
    if (pl->micro.ok) {
        e = cudaMalloc((void**)&pl->d_micro, micro_words.size() * sizeof(int32_t));
        if (e == cudaSuccess)
            e = cudaMemcpy(pl->d_micro, micro_words.data(), micro_words.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            cudaFree(pl->d_pool);
            if (pl->d_micro) cudaFree(pl->d_micro);
            return bad(PGX_ERR_CUDA, std::string("microprogram upload: ") + cudaGetErrorString(e));
        }
    }
    *out = pl;
    return PGX_OK;
}

void pgx_plan_destroy(pgx_plan* plan) {
    if (!plan) return;
    if (plan->d_pool) cudaFree(plan->d_pool);
    if (plan->d_micro) cudaFree(plan->d_micro);
    pgx::pgx_spec_destroy(plan->spec);
    if (plan->d_trace) cudaFree(plan->d_trace);
    for (GraphEntry& g : plan->graphs) cudaGraphExecDestroy(g.exec);
    for (StepSchedule& c : plan->schedules) c.release();
    if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
    if (plan->aux_stream) cudaStreamDestroy(plan->aux_stream);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    if (plan->ev_join) cudaEventDestroy(plan->ev_join);
    delete plan;
}

// true when pgx_run_batch will launch the plan-specialised kernel (no workspace: its work tables are registers)
static bool runs_specialized(const pgx_plan* plan) {
    return plan->spec && plan->n_inputs == 0 && plan->mode != PGX_MODE_STEPWISE && (plan->fused_kernel == 0 || plan->fused_kernel == 4);
}

size_t pgx_workspace_bytes(const pgx_plan* plan, int64_t B) {
    if (!plan || B <= 0) return 0;
    if (runs_specialized(plan)) return 0;
    const size_t item = plan->dtype == PGX_F64 ? 8 : 4;
    // head: room for a copy of the batch-invariant tables (32-bit-addressed step kernel); then the work tables
    const size_t head = ((size_t)plan->table_entries + 31) / 32 * 32;
    return (head + (size_t)plan->ws_entries * (size_t)pgx_batch_ld(B)) * item;
}

int pgx_plan_specialize(pgx_plan* plan) {
    if (!plan) return fail(PGX_ERR_INVALID, "null plan");
    if (plan->spec) return PGX_OK;
    if (plan->d_trace) return fail(PGX_ERR_UNSUPPORTED, "max-product plans are not specialised");
    const size_t elem = plan->dtype == PGX_F64 ? 8 : 4;
    std::vector<unsigned char> host((size_t)plan->table_entries * elem + 8);
    if (plan->table_entries > 0) PGX_CUDA(cudaMemcpy(host.data(), plan->blob, (size_t)plan->table_entries * elem, cudaMemcpyDeviceToHost));
    std::string why;
    int warps = PGX_SPEC_DEFAULT_WARPS;
    if (const char* e = std::getenv("PGX_SPEC_WARPS")) warps = std::max(1, std::atoi(e));  // tuning knob
    plan->spec = pgx::pgx_spec_build(plan->pool.data(), (int64_t)plan->pool.size(), host.data(), plan->table_entries, plan->dtype, why, warps);
    if (!plan->spec) return fail(PGX_ERR_UNSUPPORTED, "plan not specialised: " + why);
    return PGX_OK;
}

int pgx_plan_set_option(pgx_plan* plan, int32_t option, int64_t value) {
    if (!plan) return fail(PGX_ERR_INVALID, "null plan");
    switch (option) {
        case PGX_OPT_MODE:
            if (value < PGX_MODE_AUTO || value > PGX_MODE_FUSED) return fail(PGX_ERR_INVALID, "unknown mode");
            plan->mode = (int)value;
            return PGX_OK;
        case PGX_OPT_FUSED_WARPS:
            if (value < 0 || value > 16) return fail(PGX_ERR_INVALID, "fused warps must be 0..16");
            plan->fused_warps = (int)value;
            return PGX_OK;
        case PGX_OPT_USE_GRAPH:
            plan->use_graph = value ? 1 : 0;
            return PGX_OK;
        case PGX_OPT_STAGE:
        case PGX_OPT_MMA:
        case PGX_OPT_TC32:
            if (option == PGX_OPT_STAGE) {
                if (value < 0 || value > 1) return fail(PGX_ERR_INVALID, "stage must be 0 or 1");
                plan->stage = (int)value;
            } else if (option == PGX_OPT_MMA) {
                plan->mma = value ? 1 : 0;
            } else {
                plan->tc32 = value ? 1 : 0;
            }
            for (StepSchedule& c : plan->schedules) c.release();
            plan->schedules.clear();
            for (GraphEntry& g : plan->graphs) cudaGraphExecDestroy(g.exec);
            plan->graphs.clear();
            return PGX_OK;
        case PGX_OPT_STEP_KERNEL:
            if (value < 0 || value > 2) return fail(PGX_ERR_INVALID, "step kernel must be 0, 1 or 2");
            plan->step_kernel = (int)value;
            return PGX_OK;
        case PGX_OPT_FUSED_KERNEL:
            if (value < 0 || value > 4) return fail(PGX_ERR_INVALID, "fused kernel must be 0..4");
            plan->fused_kernel = (int)value;
            return PGX_OK;
        default:
            return fail(PGX_ERR_UNSUPPORTED, "unknown option");
    }
}

int pgx_plan_get_info(const pgx_plan* plan, int32_t what, int64_t* value) {
    if (!plan || !value) return fail(PGX_ERR_INVALID, "null argument");
    switch (what) {
        case PGX_INFO_N_STEPS: *value = plan->n_steps; break;
        case PGX_INFO_OUT_ELEMS: *value = plan->out_elems; break;
        case PGX_INFO_WS_ENTRIES: *value = plan->ws_entries; break;
        case PGX_INFO_LAST_LAUNCHES: *value = plan->last_launches; break;
        case PGX_INFO_LAST_MODE: *value = plan->last_mode; break;
        case PGX_INFO_N_EV: *value = plan->n_ev; break;
        case PGX_INFO_LAST_VARIANT: *value = plan->last_variant; break;
        case PGX_INFO_LAST_GRAPH: *value = plan->last_graph; break;
        case PGX_INFO_IN_ELEMS: *value = plan->in_elems; break;
        case PGX_INFO_LAST_TC_STEPS: {
            int64_t n = 0;
            if (plan->last_mode == PGX_MODE_STEPWISE && plan->last_sched >= 0 && plan->last_sched < (int)plan->schedules.size())
                for (const LaunchGroup& g : plan->schedules[plan->last_sched].groups)
                    if (g.tc) n += g.n_items;
            *value = n;
            break;
        }
        case PGX_INFO_N_LEVELS: *value = plan->micro.ok ? plan->micro.n_levels : 0; break;
        case PGX_INFO_SPECIALIZED: *value = plan->spec ? 1 : 0; break;
        case PGX_INFO_SPEC_REGS: *value = plan->spec ? pgx::pgx_spec_stats(plan->spec).regs : 0; break;
        case PGX_INFO_SPEC_SMEM: *value = plan->spec ? pgx::pgx_spec_stats(plan->spec).smem_bytes : 0; break;
        case PGX_INFO_SPEC_COMPILE_MS: *value = plan->spec ? (int64_t)(pgx::pgx_spec_stats(plan->spec).compile_s * 1e3) : 0; break;
        case PGX_INFO_SPEC_LOADS: *value = plan->spec ? pgx::pgx_spec_stats(plan->spec).loads : 0; break;
        case PGX_INFO_SPEC_FLOPS: *value = plan->spec ? pgx::pgx_spec_stats(plan->spec).flops : 0; break;
        case PGX_INFO_LAST_STAGED_STEPS: {
            int64_t n = 0;
            if (plan->last_mode == PGX_MODE_STEPWISE && plan->last_sched >= 0 && plan->last_sched < (int)plan->schedules.size())
                for (const LaunchGroup& g : plan->schedules[plan->last_sched].groups)
                    if (g.mm || g.tc) n += g.n_items;
            *value = n;
            break;
        }
        default: return fail(PGX_ERR_UNSUPPORTED, "unknown info key");
    }
    return PGX_OK;
}

}  // extern "C"

namespace {

template <typename T>
int run_typed(pgx_plan* pl, const int32_t* ev, const void* soft_v, void* out_v, void* ws_v, int64_t B, cudaStream_t st) {
    const T* soft = (const T*)soft_v;
    const int32_t* d_inputs = pl->n_inputs ? pl->d_pool + pl->inputs_off : nullptr;
    const unsigned in_blocks = pl->n_inputs ? (unsigned)((B * pl->in_elems + 255) / 256) : 0u;
    const int64_t ldb = pgx_batch_ld(B);
    const T* cst = (const T*)pl->blob;
    T* ws_all = (T*)ws_v;
    const size_t ws_off0 = ((size_t)pl->table_entries + 31) / 32 * 32;
    T* ws = ws_all + ws_off0;  // work tables start after the head reserved for the table copy
    T* out = (T*)out_v;
    int mode = pl->mode;
    if (mode == PGX_MODE_AUTO) {
        // whole-plan kernel for small models when its work tables fit shared memory, or when every step takes the
        // 64-register fast path; otherwise (hepar2-class belief-update plans, large tables) the step sequence wins
        const bool fits_smem = (size_t)pl->ws_entries * 32 * sizeof(T) + (size_t)pl->n_ev * 128 <= 226 * 1024;
        // (measured: win95pts, 1 369 work entries, fused 1.15 ms vs stepwise 1.98 ms; hepar2, 1 947 entries, 2.13 vs 1.60)
        mode = (pl->micro.ok && pl->max_joint <= 8192 && B >= 2048 && (fits_smem || pl->micro.all_fast || pl->ws_entries <= 1536)) ? PGX_MODE_FUSED
                                                                                                         : PGX_MODE_STEPWISE;
    }
    int64_t launches = 0;
    if (pl->fused_kernel == 4 && (!pl->spec || soft || pl->mode == PGX_MODE_STEPWISE))
        return fail(PGX_ERR_UNSUPPORTED, "specialised kernel requested but not available for this call (pgx_plan_specialize)");
    if (pl->spec && !soft && pl->mode != PGX_MODE_STEPWISE && (pl->fused_kernel == 0 || pl->fused_kernel == 4)) {
        // plan-specialised straight-line kernel: whenever it has been built and the caller did not ask for another one
        std::string err;
        if (pgx::pgx_spec_launch(pl->spec, pl->blob, ev, out_v, B, (void*)st, err) != 0) return fail(PGX_ERR_CUDA, err);
        mode = PGX_MODE_FUSED;
        launches = 1;
        pl->last_variant = 4;
    } else if (mode == PGX_MODE_FUSED && pl->micro.ok && pl->fused_kernel != 1) {
        // table-driven kernel; work tables in shared memory when a row of 32 evidence sets fits
        const int64_t rows = (B + 31) / 32;
        const size_t ws_smem = (size_t)pl->ws_entries * 32 * sizeof(T);
        const size_t ev_smem = (size_t)pl->n_ev * 32 * sizeof(int32_t);
        const size_t limit = 227 * 1024 - 1024;
        const bool smem = ws_smem + ev_smem <= limit && pl->fused_kernel != 3;
        const size_t dyn = (smem ? ws_smem : 0) + ev_smem;
        int G = pl->fused_warps;
        if (G <= 0) {
            if (smem) {
                // CTAs per SM by shared memory, warps per SM by registers (64 regs fast-only, ~96 otherwise)
                const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (227 * 1024) / (dyn + 1024)));
                const int warps_per_sm = pl->micro.all_fast ? 32 : 20;
                G = warps_per_sm / per_sm;
                if (G < 4) G = 4;
                // every level is cut into 16 chunks of equal model cost (build_micro): with 16 warps each warp takes one
                // chunk per level and all arrive at the level barrier together. Measured on alarm (B = 131 072, fast-only,
                // 64 registers): 3 CTAs x 10 warps 0.584 ms, 2 CTAs x 16 warps 0.493 ms (10 chunks for 10 warps: 0.551).
                if (pl->micro.all_fast && per_sm >= 2) G = 16;
            } else {
                G = (int)((148 * 16 + rows - 1) / rows);
                if (G < 4) G = 4;
            }
        }
        if (G < 1) G = 1;
        if (G > 16) G = 16;
#define PGX_LAUNCH_FUSED2(SM, FO)                                                                                     \
    k_plan_fused2<T, SM, FO><<<(unsigned)rows, 32 * G, dyn, st>>>(pl->d_micro, cst, ws, ev, pl->d_pool + pl->ev_card_off, out, \
                                                                  pl->n_ev, (int)pl->ws_entries, B, ldb, soft, d_inputs)
        if (smem && pl->micro.all_fast) {
            PGX_CUDA(cudaFuncSetAttribute(k_plan_fused2<T, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            PGX_LAUNCH_FUSED2(true, true);
        } else if (smem) {
            PGX_CUDA(cudaFuncSetAttribute(k_plan_fused2<T, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
            PGX_LAUNCH_FUSED2(true, false);
        } else if (pl->micro.all_fast) {
            PGX_LAUNCH_FUSED2(false, true);
        } else {
            PGX_LAUNCH_FUSED2(false, false);
        }
#undef PGX_LAUNCH_FUSED2
        PGX_CUDA(cudaGetLastError());
        launches = 1;
        pl->last_variant = smem ? 2 : 3;
    } else if (mode == PGX_MODE_FUSED) {
        pl->last_variant = 1;
        int G = pl->fused_warps;
        const int64_t rows = (B + 31) / 32;
        if (G <= 0) {
            G = (int)((148 * 16 + rows - 1) / rows);
            if (G < 1) G = 1;
            if (G > 16) G = 16;
        }
        if (G > 16) {
            G = 16;
        }
        dim3 block(32, G);
        if (pl->n_inputs) k_scatter_inputs<T><<<in_blocks, 256, 0, st>>>(d_inputs, soft, ws, B, ldb);
        if (pl->max_ops <= 4)
            k_plan_fused<T, 4><<<(unsigned)rows, block, 0, st>>>(pl->d_pool, cst, ws, ev, out, B, ldb);
        else if (pl->max_ops <= 8)
            k_plan_fused<T, 8><<<(unsigned)rows, block, 0, st>>>(pl->d_pool, cst, ws, ev, out, B, ldb);
        else
            k_plan_fused<T, MAX_OPS><<<(unsigned)rows, block, 0, st>>>(pl->d_pool, cst, ws, ev, out, B, ldb);
        PGX_CUDA(cudaGetLastError());
        launches = 1;
    } else {
        const int64_t bt = ldb < 32 ? ldb : 32;
        const int bt_log2 = ilog2_floor(bt);
        const int64_t b_tiles = (B + bt - 1) / bt;
        if (b_tiles > 65535) return fail(PGX_ERR_UNSUPPORTED, "batch too large for one stepwise launch (max 2,097,120)");
        const int per_block = 256 >> bt_log2;
        // launch schedule for this batch size: tile-eligible steps of one dependency level share a launch
        StepSchedule* sched = nullptr;
        for (StepSchedule& c : pl->schedules)
            if (c.B == B && c.step_kernel == pl->step_kernel + 4 * pl->batch_levels + 32 * pl->stage + 128 * pl->mma + 256 * pl->tc32 &&
                c.dtype_size == (int)sizeof(T))
                sched = &c;
        if (!sched) {
            StepSchedule ns;
            ns.B = B;
            ns.step_kernel = pl->step_kernel + 4 * pl->batch_levels + 32 * pl->stage + 128 * pl->mma + 256 * pl->tc32;
            const int64_t tile_b_tiles = b_tiles;
            ns.dtype_size = (int)sizeof(T);
            std::vector<TileItem> items;
            const int o_per_warp = 32 >> bt_log2;
            std::vector<TileItem> pending;  // items of the open tile group of the current level
            int cur_level = -1;
            auto flush = [&](LaunchGroup& g) {
                if (g.n_items > 0) {
                    g.first_item = (int)items.size();
                    items.insert(items.end(), pending.begin(), pending.end());
                    pending.clear();
                    g.level = cur_level;
                    ns.groups.push_back(g);
                }
                g = LaunchGroup();
            };
            // per dependency level: one group of tile steps, generic steps alone
            LaunchGroup cur;
            // ... and one group of steps for the matrix-product tile kernel (pgx_mm.cu)
            LaunchGroup cur_mm, cur_tc;
            std::vector<MMItem> pending_tc;
            const int b_chunks = (int)((B + 127) / 128);
            std::vector<MMItem> mm_items, pending_mm;
            std::vector<double> pending_mm_cost;
            std::vector<int32_t> mm_tabs;
            const bool stage_on = pl->stage && pl->step_kernel == 0 && bt_log2 == 5 &&
                                  ws_off0 + (size_t)pl->ws_entries * (size_t)ldb < (1ULL << 32) && pl->ws_entries < (1LL << 31);
            auto flush_level = [&]() {
                flush(cur);
                if (cur_mm.n_items > 0) {
                    // split every step's tiles into CTA-sized runs. One CTA per SM is resident (the stage ring takes the
                    // shared memory), so the launch gets a whole number of waves: `waves` x resident CTAs dealt to the steps in
                    // proportion to their model cost, a run never shorter than ~8k model cycles (pipeline prologue).
                    double total = 0;
                    for (size_t i = 0; i < pending_mm.size(); ++i) total += pending_mm_cost[i] * (double)b_tiles;
                    const double slots = 148.0 * MM_CTAS_PER_SM;  // CTAs resident at once
                    // measured (diabetes / munin, ms): one wave of long runs 53.2 / 32.9; runs of ~12k model cycles (several
                    // waves, dynamically scheduled: the cost model's errors even out) 50.0 / 31.9; shorter runs 50.5-50.8
                    double wave_cycles = 12000.0, min_cycles = 8000.0;
                    if (const char* e = std::getenv("PGX_MM_WAVE_CYCLES")) wave_cycles = std::atof(e);  // tuning knobs
                    if (const char* e = std::getenv("PGX_MM_MIN_CYCLES")) min_cycles = std::atof(e);
                    int waves = (int)(total / (slots * wave_cycles));
                    waves = waves < 1 ? 1 : (waves > 48 ? 48 : waves);
                    const double target = std::max(min_cycles, total / (slots * waves));
                    int nb = 0;
                    for (size_t i = 0; i < pending_mm.size(); ++i) {
                        MMItem& mi = pending_mm[i];
                        const int64_t tiles = (int64_t)mi.n_tiles * b_tiles;
                        const double per_tile = pending_mm_cost[i] / (double)mi.n_tiles;
                        int64_t want = (int64_t)(per_tile * (double)tiles / target + 0.5);  // CTAs for this step
                        want = std::max<int64_t>(1, std::min<int64_t>(want, tiles));
                        const int64_t tpc = (tiles + want - 1) / want;
                        mi.tiles_per_cta = (int32_t)tpc;
                        mi.n_ctas = (int32_t)((tiles + tpc - 1) / tpc);
                        mi.blk_begin = nb;
                        nb += mi.n_ctas;
                    }
                    cur_mm.n_blocks = nb;
                    cur_mm.first_item = (int)mm_items.size();
                    mm_items.insert(mm_items.end(), pending_mm.begin(), pending_mm.end());
                    pending_mm.clear();
                    pending_mm_cost.clear();
                    cur_mm.level = cur_level;
                    ns.groups.push_back(cur_mm);
                }
                cur_mm = LaunchGroup();
                if (cur_tc.n_items > 0) {
                    int nb = 0;
                    for (MMItem& mi : pending_tc) {
                        const int npc = mi.N < 16 ? mi.N : 16;  // n per CTA: the CPT slice is staged once per CTA
                        mi.tiles_per_cta = npc;
                        mi.n_ctas = mi.Z * b_chunks * ((mi.N + npc - 1) / npc);
                        mi.blk_begin = nb;
                        nb += mi.n_ctas;
                    }
                    cur_tc.n_blocks = nb;
                    cur_tc.first_item = (int)mm_items.size();
                    mm_items.insert(mm_items.end(), pending_tc.begin(), pending_tc.end());
                    pending_tc.clear();
                    cur_tc.level = cur_level;
                    ns.groups.push_back(cur_tc);
                }
                cur_tc = LaunchGroup();
            };
            for (size_t si = 0; si < pl->steps.size(); ++si) {
                const StepInfo& s = pl->steps[si];
                const int64_t stab_words = s.sum_size * s.n_ops;
                int n_divisors = 0;
                {
                    const int32_t* r0 = pl->pool.data() + s.rec_off;
                    const int opw0 = OP_FIXED + r0[0] + r0[1];
                    for (int k = 0; k < s.n_ops; ++k)
                        if (r0[STEP_FIXED + r0[0] + r0[1] + k * opw0] & 0x100) ++n_divisors;
                }
                // (the 32-bit tile kernel keeps at most two divisor rows in registers; a step of divisors only is generic)
                const bool tile_ok = s.n_ops <= 8 && stab_words <= 8192 && pl->step_kernel != 1 &&  // 0 tile32, 2 tile64
                                     n_divisors <= 2 && n_divisors < s.n_ops;
                const int32_t* srec = pl->pool.data() + s.rec_off;
                if (s.level != cur_level || !pl->batch_levels) flush_level();
                cur_level = s.level;
                if (stage_on) {
                    MMChoice ch;
                    if (mm_pick(srec, sizeof(T), pl->mma != 0, ldb, ch) &&
                        (int64_t)ch.item.n_tiles * b_tiles < (1LL << 30) && mm_tabs.size() + ch.tabs.size() < (1u << 30)) {
                        ch.item.tab = (int32_t)mm_tabs.size();
                        mm_tabs.insert(mm_tabs.end(), ch.tabs.begin(), ch.tabs.end());
                        if (sizeof(T) == 4 && pl->tc32 && B >= 128 && tc32_eligible(ch.item) &&
                            (int64_t)ch.item.Z * b_chunks * ch.item.N < (1LL << 30)) {
                            // fp32 mode, CPT x message: the tcgen05 tensor cores (TF32x3, TMEM accumulators)
                            pending_tc.push_back(ch.item);
                            cur_tc.tc = true;
                            cur_tc.step_ids.push_back((int)si);
                            cur_tc.n_items += 1;
                            cur_tc.smem = std::max(cur_tc.smem, tc32_smem_bytes(ch.item));
                            continue;
                        }
                        pending_mm.push_back(ch.item);
                        pending_mm_cost.push_back(ch.cost);
                        cur_mm.mm = true;
                        cur_mm.step_ids.push_back((int)si);
                        cur_mm.n_items += 1;
                        cur_mm.smem = std::max(cur_mm.smem, ch.smem);
                        continue;
                    }
                }
                if (!tile_ok) {
                    flush_level();
                    LaunchGroup g;
                    g.generic_step = (int)si;
                    g.n_items = 1;
                    g.level = s.level;
                    g.step_ids.push_back((int)si);
                    ns.groups.push_back(g);
                    continue;
                }
                static const int knob_btb = std::getenv("PGX_TILE_BTB") ? std::atoi(std::getenv("PGX_TILE_BTB")) : 4;        // tuning knobs
                static const int knob_to = std::getenv("PGX_TILE_TO_CAP") ? std::atoi(std::getenv("PGX_TILE_TO_CAP")) : 1024;  // measured 256..1024 x btb 2..8: within 2 %, 1024 x 4 best on munin/pathfinder
                int btb = (int)(tile_b_tiles < knob_btb ? tile_b_tiles : knob_btb);
                int64_t b_blocks = (tile_b_tiles + btb - 1) / btb;
                int64_t TO = (s.out_size * b_blocks) / (148 * 4);  // aim at >= 4 CTAs per SM when there is work
                if (TO < 8 * o_per_warp) {
                    // little work along the output: one evidence tile per CTA, one entry per warp
                    btb = 1;
                    b_blocks = tile_b_tiles;
                    TO = 8 * o_per_warp;
                }
                if (TO > knob_to) TO = knob_to;
                if (TO * s.n_ops > 4 * knob_to) TO = 4 * knob_to / s.n_ops;
                TO = (TO + o_per_warp - 1) / o_per_warp * o_per_warp;
                if (TO > s.out_size) TO = (s.out_size + o_per_warp - 1) / o_per_warp * o_per_warp;
                LaunchGroup& cg = cur;
                const int64_t n_blocks = ((s.out_size + TO - 1) / TO) * b_blocks;
                if (cg.n_items > 0 && (int64_t)cg.n_blocks + n_blocks > (1LL << 30)) flush(cg);
                pending.push_back(TileItem{s.rec_off, s.rec_len, (int32_t)TO, btb, (int32_t)b_blocks, (int32_t)cg.n_blocks});
                cg.step_ids.push_back((int)si);
                cg.n_items += 1;
                cg.n_blocks += (int)n_blocks;
                cg.max_k = std::max(cg.max_k, s.n_ops);
                cg.smem = std::max(cg.smem, (size_t)(((s.rec_len + 3) & ~3) + TO * s.n_ops + stab_words) * sizeof(int32_t));
            }
            flush_level();
            if (!items.empty()) {
                PGX_CUDA(cudaMalloc((void**)&ns.d_items, items.size() * sizeof(TileItem)));
                PGX_CUDA(cudaMemcpy(ns.d_items, items.data(), items.size() * sizeof(TileItem), cudaMemcpyHostToDevice));
            }
            if (!mm_items.empty()) {
                PGX_CUDA(cudaMalloc((void**)&ns.d_mm_items, mm_items.size() * sizeof(MMItem)));
                PGX_CUDA(cudaMemcpy(ns.d_mm_items, mm_items.data(), mm_items.size() * sizeof(MMItem), cudaMemcpyHostToDevice));
                PGX_CUDA(cudaMalloc((void**)&ns.d_mm_tabs, std::max<size_t>(1, mm_tabs.size()) * sizeof(int32_t)));
                PGX_CUDA(cudaMemcpy(ns.d_mm_tabs, mm_tabs.data(), mm_tabs.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            }
            if (pl->schedules.size() >= 8) {
                pl->schedules.front().release();
                pl->schedules.erase(pl->schedules.begin());
                for (GraphEntry& g : pl->graphs) cudaGraphExecDestroy(g.exec);  // graphs reference the item tables
                pl->graphs.clear();
            }
            pl->schedules.push_back(ns);
            sched = &pl->schedules.back();
        }
        pl->last_sched = (int)(sched - pl->schedules.data());
        // 32-bit addressing needs every element index (table copy + work tables) below 2^32
        const bool idx32 = pl->step_kernel == 0 && ws_off0 + (size_t)pl->ws_entries * (size_t)ldb < (1ULL << 32) &&
                           pl->ws_entries < (1LL << 31);
        if (pl->level_streams && !pl->aux_stream) {
            PGX_CUDA(cudaStreamCreateWithFlags(&pl->aux_stream, cudaStreamNonBlocking));
            PGX_CUDA(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
            PGX_CUDA(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
        }
        const int b_chunks_rt = (int)((B + 127) / 128);
        auto enqueue = [&](cudaStream_t qs) -> int {
            int n = 0;
            cudaEvent_t* evs = pl->prof_events;
            int ev_idx = 0;
            if (idx32) cudaMemcpyAsync(ws_all, cst, (size_t)pl->table_entries * sizeof(T), cudaMemcpyDeviceToDevice, qs);
            if (pl->n_inputs) {
                k_scatter_inputs<T><<<in_blocks, 256, 0, qs>>>(d_inputs, soft, ws, B, ldb);
                ++n;
            }
            static const int prof_sync = std::getenv("PGX_PROFILE_SYNC") ? std::atoi(std::getenv("PGX_PROFILE_SYNC")) : 0;
            static const int knob_streams = std::getenv("PGX_LEVEL_STREAMS") ? std::atoi(std::getenv("PGX_LEVEL_STREAMS")) : 1;
            const bool fork_levels = pl->level_streams && knob_streams && !evs && pl->aux_stream && pl->ev_fork && pl->ev_join;
            const cudaStream_t main_q = qs;
            int lvl = -1 << 30, k_in_level = 0;
            bool forked = false;
            const size_t n_groups = sched->groups.size();
            for (size_t gi = 0; gi < n_groups; ++gi) {
                const LaunchGroup& g = sched->groups[gi];
                if (g.level != lvl) {
                    if (forked) {  // join: the next level needs everything of this one
                        cudaEventRecord(pl->ev_join, pl->aux_stream);
                        cudaStreamWaitEvent(main_q, pl->ev_join, 0);
                        forked = false;
                    }
                    lvl = g.level;
                    k_in_level = 0;
                    if (fork_levels && gi + 1 < n_groups && sched->groups[gi + 1].level == lvl) {
                        cudaEventRecord(pl->ev_fork, main_q);  // = the previous level is complete
                        cudaStreamWaitEvent(pl->aux_stream, pl->ev_fork, 0);
                        forked = true;
                    }
                }
                qs = (forked && (k_in_level & 1)) ? pl->aux_stream : main_q;
                ++k_in_level;
                if (evs && prof_sync) {  // tracing aid: let the previous launch (and its write-backs) drain first
                    cudaStreamSynchronize(qs);
                    if (prof_sync > 1) {  // ... and push its dirty lines out of L2 by reading the table head
                        cudaMemsetAsync(ws_all, 0, 0, qs);
                    }
                    cudaEventRecord(evs[ev_idx], qs);
                }
                if (g.generic_step >= 0) {
                    const StepInfo& s = pl->steps[g.generic_step];
                    // generic kernel: consecutive entries per thread, as many as leave >= ~4 waves of CTAs on 148 SMs
                    int64_t opt = (s.out_size * b_tiles) / (per_block * 148LL * 8 * 4);
                    opt = opt < 1 ? 1 : (opt > 16 ? 16 : opt);
                    const int64_t runs = (s.out_size + opt - 1) / opt;
                    dim3 grid((unsigned)((runs + per_block - 1) / per_block), (unsigned)b_tiles);
                    const size_t smem = (size_t)s.rec_len * sizeof(int32_t);
#define PGX_LAUNCH_STEP(MK)                                                                                      \
    k_contract_step<T, MK><<<grid, 256, smem, qs>>>(pl->d_pool, s.rec_off, s.rec_len, pl->ev_card_off, cst, ws, ev, \
                                                    pl->n_ev, B, ldb, bt_log2, (int)opt)
                    if (s.n_ops <= 2)
                        PGX_LAUNCH_STEP(2);
                    else if (s.n_ops <= 4)
                        PGX_LAUNCH_STEP(4);
                    else if (s.n_ops <= 8)
                        PGX_LAUNCH_STEP(8);
                    else
                        PGX_LAUNCH_STEP(MAX_OPS);
#undef PGX_LAUNCH_STEP
                } else if (g.tc) {
                    tc32_launch(sched->d_mm_items + g.first_item, g.n_items, g.n_blocks, g.smem, sched->d_mm_tabs, ws_all,
                                (uint32_t)ws_off0, B, (uint32_t)ldb, b_chunks_rt, qs);
                } else if (g.mm) {
                    mm_launch(sizeof(T), sched->d_mm_items + g.first_item, g.n_items, g.n_blocks, g.smem, sched->d_mm_tabs, ws_all,
                              (uint32_t)ws_off0, B, (uint32_t)ldb, (int)b_tiles, qs);
                } else {
                    const TileItem* d_it = sched->d_items + g.first_item;
#define PGX_LAUNCH_TILE(MK)                                                                                            \
    k_contract_tile<T, MK><<<(unsigned)g.n_blocks, 256, g.smem, qs>>>(pl->d_pool, d_it, g.n_items, pl->ev_card_off, cst, \
                                                                      ws, ws, ev, pl->n_ev, B, ldb, bt_log2)
#define PGX_LAUNCH_TILE32(MK)                                                                                          \
    k_contract_tile32<T, MK><<<(unsigned)g.n_blocks, 256, g.smem, qs>>>(pl->d_pool, d_it, g.n_items, pl->ev_card_off,   \
                                                                        ws_all, ws_all, (uint32_t)ws_off0, ev, pl->n_ev, \
                                                                        B, (uint32_t)ldb, bt_log2)
                    if (idx32) {
                        if (g.max_k <= 2) PGX_LAUNCH_TILE32(2);
                        else if (g.max_k <= 4) PGX_LAUNCH_TILE32(4);
                        else if (g.max_k <= 6) PGX_LAUNCH_TILE32(6);  // 48 registers, 40 warps per SM (pathfinder's hub belief: K = 5)
                        else PGX_LAUNCH_TILE32(8);
                    } else if (g.max_k <= 2) {
                        PGX_LAUNCH_TILE(2);
                    } else if (g.max_k <= 4) {
                        PGX_LAUNCH_TILE(4);
                    } else {
                        PGX_LAUNCH_TILE(8);
                    }
#undef PGX_LAUNCH_TILE32
#undef PGX_LAUNCH_TILE
                }
                ++n;
                if (evs) cudaEventRecord(evs[++ev_idx], qs);
            }
            if (forked) {
                cudaEventRecord(pl->ev_join, pl->aux_stream);
                cudaStreamWaitEvent(main_q, pl->ev_join, 0);
            }
            qs = main_q;
            if (pl->n_segs > 0) {
                dim3 grid((unsigned)((B + 127) / 128), (unsigned)pl->n_segs);
                k_emit<T><<<grid, 128, 0, qs>>>(pl->d_pool, pl->segs_off, ws, out, pl->out_elems, B, ldb);
                ++n;
            }
            return n;
        };
        if (pl->n_segs > 65535) return fail(PGX_ERR_UNSUPPORTED, "too many output segments");
        bool replayed = false;
        if (pl->use_graph && pl->n_steps >= 8 && !pl->prof_events) {
            // replay the launch sequence as a CUDA graph (captured once per argument tuple)
            GraphEntry* hit = nullptr;
            for (GraphEntry& g : pl->graphs)
                if (g.B == B && g.ev == (const void*)ev && g.soft == soft_v && g.out == out_v && g.ws == ws_v && g.step_kernel == pl->step_kernel &&
                    g.dtype_size == (int)sizeof(T))
                    hit = &g;
            // capture + instantiate cost as much as dozens of direct passes (munin: thousands of nodes), so a tuple is
            // captured only when it shows up a second time among the recent misses: callers that hand over fresh buffers on
            // every call (out[lo:hi] slices, new evidence tensors) then never pay for a graph they will not replay
            const GraphEntry key{B, (const void*)ev, soft_v, out_v, ws_v, pl->step_kernel, (int)sizeof(T), 0, nullptr};
            bool seen_before = false;
            for (const GraphEntry& c : pl->graph_candidates)
                if (c.B == key.B && c.ev == key.ev && c.soft == key.soft && c.out == key.out && c.ws == key.ws &&
                    c.step_kernel == key.step_kernel && c.dtype_size == key.dtype_size)
                    seen_before = true;
            if (!hit && !seen_before) {  // remember the last few misses (callers often alternate two or three buffer sets)
                if (pl->graph_candidates.size() >= 8) pl->graph_candidates.erase(pl->graph_candidates.begin());
                pl->graph_candidates.push_back(key);
            }
            if (!hit && seen_before) {
                if (!pl->cap_stream) PGX_CUDA(cudaStreamCreateWithFlags(&pl->cap_stream, cudaStreamNonBlocking));
                cudaGraph_t graph = nullptr;
                PGX_CUDA(cudaStreamBeginCapture(pl->cap_stream, cudaStreamCaptureModeThreadLocal));
                const int n = enqueue(pl->cap_stream);
                cudaError_t ce = cudaStreamEndCapture(pl->cap_stream, &graph);
                if (ce == cudaSuccess && graph) {
                    cudaGraphExec_t exec = nullptr;
                    ce = cudaGraphInstantiate(&exec, graph, 0);
                    cudaGraphDestroy(graph);
                    if (ce == cudaSuccess) {
                        if (pl->graphs.size() >= 8) {
                            // the evicted executable may still be running on another stream of the caller
                            cudaDeviceSynchronize();
                            cudaGraphExecDestroy(pl->graphs.front().exec);
                            pl->graphs.erase(pl->graphs.begin());
                        }
                        pl->graphs.push_back(GraphEntry{B, (const void*)ev, soft_v, out_v, ws_v, pl->step_kernel, (int)sizeof(T), n, exec});
                        hit = &pl->graphs.back();
                    }
                }
                if (!hit) (void)cudaGetLastError();  // capture failed: fall through to direct launches
            }
            if (hit) {
                PGX_CUDA(cudaGraphLaunch(hit->exec, st));
                launches = hit->n_launches;
                replayed = true;
            }
        }
        if (!replayed) {
            launches = enqueue(st);
            PGX_CUDA(cudaGetLastError());
        }
        pl->last_graph = replayed ? 1 : 0;
    }
    if (mode != PGX_MODE_FUSED) pl->last_variant = 0;
    pl->last_launches = launches;
    pl->last_mode = mode;
    return PGX_OK;
}

}  // namespace

extern "C" {

int pgx_run_batch(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                  int64_t B, void* stream) {
    return pgx_run_batch_soft(plan, ev_states, nullptr, out, workspace, workspace_bytes, B, stream);
}

int pgx_run_batch_multi(int32_t n, pgx_plan* const* plans, const int32_t* const* ev_states, void* const* outs,
                        void* const* workspaces, const size_t* workspace_bytes, const int64_t* B, void* stream) {
    if (n < 0 || (n > 0 && (!plans || !ev_states || !outs || !workspaces || !workspace_bytes || !B)))
        return fail(PGX_ERR_INVALID, "null argument");
    for (int32_t i = 0; i < n; ++i) {
        const int rc = pgx_run_batch(plans[i], ev_states[i], outs[i], workspaces[i], workspace_bytes[i], B[i], stream);
        if (rc != PGX_OK) {
            g_err = "job " + std::to_string(i) + ": " + g_err;
            return rc;
        }
    }
    return PGX_OK;
}

int pgx_run_batch_soft(pgx_plan* plan, const int32_t* ev_states, const void* soft, void* out, void* workspace,
                       size_t workspace_bytes, int64_t B, void* stream) {
    if (!plan) return fail(PGX_ERR_INVALID, "null plan");
    if (plan->n_inputs > 0 && !soft) return fail(PGX_ERR_INVALID, "plan has soft-evidence input tables but `soft` is null");
    if (plan->n_inputs == 0 && soft) return fail(PGX_ERR_INVALID, "plan has no soft-evidence input tables");
    if (B <= 0) return fail(PGX_ERR_INVALID, "batch must be positive");
    if (plan->n_ev > 0 && !ev_states) return fail(PGX_ERR_INVALID, "plan has evidence slots but ev_states is null");
    if (!out && plan->out_elems > 0) return fail(PGX_ERR_INVALID, "null output");
    if ((!workspace && pgx_workspace_bytes(plan, B) > 0) || workspace_bytes < pgx_workspace_bytes(plan, B))
        return fail(PGX_ERR_WORKSPACE, "workspace too small: need " + std::to_string(pgx_workspace_bytes(plan, B)) + " bytes");
    cudaStream_t st = (cudaStream_t)stream;
    if (plan->dtype == PGX_F64) return run_typed<double>(plan, ev_states, soft, out, workspace, B, st);
    return run_typed<float>(plan, ev_states, soft, out, workspace, B, st);
}

int pgx_plan_set_trace(pgx_plan* plan, const int32_t* trace, int64_t n_words) {
    if (!plan || !trace || n_words < 2) return fail(PGX_ERR_INVALID, "bad argument");
    const int n_cliques = trace[0], n_cols = trace[1];
    if (n_cliques < 1 || n_cols < 1) return fail(PGX_ERR_INVALID, "empty traceback descriptor");
    int64_t at = 2;
    std::vector<char> seen((size_t)n_cols, 0);
    for (int c = 0; c < n_cliques; ++c) {
        if (at + 3 > n_words) return fail(PGX_ERR_INVALID, "traceback descriptor truncated");
        const int64_t off = ld_i64(trace + at);
        const int n_ax = trace[at + 2];
        if (n_ax < 0 || n_ax > MAX_AXES || at + 3 + 4LL * n_ax > n_words) return fail(PGX_ERR_INVALID, "traceback descriptor truncated");
        int64_t hi = off;
        for (int a = 0; a < n_ax; ++a) {
            const int32_t* ax = trace + at + 3 + 4 * a;
            if (ax[0] < 0 || ax[0] >= n_cols || ax[1] < 1 || ax[2] < 0) return fail(PGX_ERR_INVALID, "bad traceback axis");
            if (ax[3]) {
                if (seen[ax[0]]) return fail(PGX_ERR_INVALID, "traceback column assigned twice");
                seen[ax[0]] = 1;
            } else if (!seen[ax[0]]) {
                return fail(PGX_ERR_INVALID, "traceback reads a column before it is assigned");
            }
            hi += (int64_t)(ax[1] - 1) * ax[2];
        }
        if (off < 0 || hi >= plan->ws_entries) return fail(PGX_ERR_BOUNDS, "traceback table outside workspace");
        at += 3 + 4 * n_ax;
    }
    for (int i = 0; i < n_cols; ++i)
        if (!seen[i]) return fail(PGX_ERR_INVALID, "traceback leaves a column unassigned");
    if (plan->d_trace) cudaFree(plan->d_trace);
    plan->d_trace = nullptr;
    PGX_CUDA(cudaMalloc((void**)&plan->d_trace, (size_t)n_words * sizeof(int32_t)));
    PGX_CUDA(cudaMemcpy(plan->d_trace, trace, (size_t)n_words * sizeof(int32_t), cudaMemcpyHostToDevice));
    plan->trace_cols = n_cols;
    return PGX_OK;
}

int pgx_run_batch_mpe(pgx_plan* plan, const int32_t* ev_states, const void* soft, int32_t* assign, void* workspace,
                      size_t workspace_bytes, int64_t B, void* stream) {
    if (!plan || !assign) return fail(PGX_ERR_INVALID, "null argument");
    if (!plan->d_trace) return fail(PGX_ERR_INVALID, "plan has no traceback descriptor (pgx_plan_set_trace)");
    if (plan->out_elems != 0) return fail(PGX_ERR_INVALID, "a max-product plan has no output segments");
    const int saved_mode = plan->mode;
    plan->mode = PGX_MODE_STEPWISE;  // the traceback reads the beliefs from the global workspace
    const int rc = pgx_run_batch_soft(plan, ev_states, soft, nullptr, workspace, workspace_bytes, B, stream);
    plan->mode = saved_mode;
    if (rc != PGX_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ldb = pgx_batch_ld(B);
    const size_t head = ((size_t)plan->table_entries + 31) / 32 * 32;
    const unsigned grid = (unsigned)((B + 127) / 128);
    if (plan->dtype == PGX_F64)
        k_mpe_traceback<double><<<grid, 128, 0, st>>>(plan->d_trace, (const double*)workspace + head, assign, B, ldb);
    else
        k_mpe_traceback<float><<<grid, 128, 0, st>>>(plan->d_trace, (const float*)workspace + head, assign, B, ldb);
    PGX_CUDA(cudaGetLastError());
    plan->last_launches += 1;
    return PGX_OK;
}

// Shared body of the two tracing entry points: a stepwise pass with one CUDA event per launch.
static int profile_pass(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                        int64_t B, void* stream, bool per_step, std::vector<float>& ms) {
    const int saved_mode = plan->mode, saved_batch = plan->batch_levels;
    std::vector<cudaEvent_t> evs((size_t)plan->n_steps + 2, nullptr);
    int rc = PGX_OK;
    for (auto& e : evs)
        if (cudaEventCreate(&e) != cudaSuccess) rc = fail(PGX_ERR_CUDA, "cudaEventCreate failed");
    if (rc == PGX_OK) {
        plan->mode = PGX_MODE_STEPWISE;
        if (per_step) plan->batch_levels = 0;  // one launch per step so that every step gets its own event pair
        plan->prof_events = evs.data();
        if (cudaEventRecord(evs[0], (cudaStream_t)stream) != cudaSuccess) rc = fail(PGX_ERR_CUDA, "cudaEventRecord failed");
        if (rc == PGX_OK) rc = pgx_run_batch(plan, ev_states, out, workspace, workspace_bytes, B, stream);
        plan->prof_events = nullptr;
        plan->mode = saved_mode;
        plan->batch_levels = saved_batch;
    }
    if (rc == PGX_OK && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess)
        rc = fail(PGX_ERR_CUDA, "cudaStreamSynchronize failed");
    if (rc == PGX_OK) {
        const size_t n = plan->last_sched >= 0 ? plan->schedules[plan->last_sched].groups.size() : 0;
        ms.assign(n, 0.f);
        for (size_t i = 0; i < n && i + 1 < evs.size(); ++i)
            if (cudaEventElapsedTime(&ms[i], evs[i], evs[i + 1]) != cudaSuccess) rc = fail(PGX_ERR_CUDA, "cudaEventElapsedTime failed");
    }
    for (auto& e : evs)
        if (e) cudaEventDestroy(e);
    return rc;
}

int pgx_profile_steps(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                      int64_t B, void* stream, float* step_ms, int32_t n_steps) {
    if (!plan || !step_ms) return fail(PGX_ERR_INVALID, "null argument");
    if (n_steps < plan->n_steps) return fail(PGX_ERR_INVALID, "step_ms too short");
    std::vector<float> ms;
    const int rc = profile_pass(plan, ev_states, out, workspace, workspace_bytes, B, stream, true, ms);
    if (rc != PGX_OK) return rc;
    // with one launch per step the launch groups are the steps, in plan order
    const StepSchedule& sc = plan->schedules[plan->last_sched];
    for (int i = 0; i < plan->n_steps; ++i) step_ms[i] = 0.f;
    for (size_t g = 0; g < sc.groups.size(); ++g)
        for (int si : sc.groups[g].step_ids) step_ms[si] += ms[g];
    return PGX_OK;
}

int pgx_profile_launches(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                         int64_t B, void* stream, float* launch_ms, int32_t cap_launches, int32_t* step_launch,
                         int32_t n_steps, int32_t* n_launches) {
    if (!plan || !launch_ms || !step_launch || !n_launches) return fail(PGX_ERR_INVALID, "null argument");
    if (n_steps < plan->n_steps) return fail(PGX_ERR_INVALID, "step_launch too short");
    std::vector<float> ms;
    const int rc = profile_pass(plan, ev_states, out, workspace, workspace_bytes, B, stream, false, ms);
    if (rc != PGX_OK) return rc;
    const StepSchedule& sc = plan->schedules[plan->last_sched];
    if ((int64_t)sc.groups.size() > cap_launches) return fail(PGX_ERR_INVALID, "launch_ms too short");
    *n_launches = (int32_t)sc.groups.size();
    for (size_t g = 0; g < sc.groups.size(); ++g) {
        launch_ms[g] = ms[g];
        for (int si : sc.groups[g].step_ids) step_launch[si] = (int32_t)g;
    }
    return PGX_OK;
}

int pgx_mm_pick(const int32_t* step_record, int32_t item_bytes, int32_t allow_mma, int32_t* fields, int32_t* tabs,
                int64_t tabs_cap, int64_t* n_tabs) {
    if (!step_record || !fields || !n_tabs || (item_bytes != 4 && item_bytes != 8)) return fail(PGX_ERR_INVALID, "bad argument");
    pgx::MMChoice ch;
    const bool ok = pgx::mm_pick(step_record, (size_t)item_bytes, allow_mma != 0, 1, ch);
    for (int i = 0; i < 24; ++i) fields[i] = 0;
    *n_tabs = 0;
    if (!ok) return PGX_OK;
    const pgx::MMItem& it = ch.item;
    const int32_t f[24] = {1, it.M, it.N, it.Z, it.K, it.lgTX, it.lgTY, it.TZ, it.lgKC, it.ntx, it.nty, it.ntz, it.n_chunks,
                           it.n_stages, it.stage_elems, it.q_off, it.p_const, (int32_t)it.p_base, (int32_t)it.q_base,
                           (int32_t)it.o_base, it.n_active, it.use_mma, (int32_t)ch.smem, it.n_tiles};
    for (int i = 0; i < 24; ++i) fields[i] = f[i];
    *n_tabs = (int64_t)ch.tabs.size();
    if (tabs && tabs_cap >= (int64_t)ch.tabs.size()) std::memcpy(tabs, ch.tabs.data(), ch.tabs.size() * sizeof(int32_t));
    return PGX_OK;
}

int pgx_evidence_reduce(int32_t dtype, const void* table, int64_t table_entries, int32_t n_free,
                        const int32_t* free_dims, const int32_t* free_strides, int32_t n_ev, const int32_t* ev_slots,
                        const int32_t* ev_strides, const int32_t* ev_cards, const int32_t* ev_states, int32_t ev_row_len,
                        void* dst, int64_t B, int64_t ldb, void* stream) {
    if (!table || !dst || B <= 0 || n_free < 0 || n_ev < 0 || n_free > MAX_AXES || n_ev > MAX_AXES)
        return fail(PGX_ERR_INVALID, "bad argument");
    if (n_ev > 0 && !ev_states) return fail(PGX_ERR_INVALID, "null evidence");
    if (ldb < B) return fail(PGX_ERR_INVALID, "ldb < B");
    int64_t n = 1, hi = 0;
    for (int a = 0; a < n_free; ++a) {
        if (free_dims[a] < 1 || free_strides[a] < 0) return fail(PGX_ERR_INVALID, "bad free axis");
        n *= free_dims[a];
        hi += (int64_t)(free_dims[a] - 1) * free_strides[a];
    }
    for (int j = 0; j < n_ev; ++j) {
        if (ev_cards[j] < 1 || ev_strides[j] < 0 || ev_slots[j] < 0 || ev_slots[j] >= ev_row_len)
            return fail(PGX_ERR_INVALID, "bad evidence axis");
        hi += (int64_t)(ev_cards[j] - 1) * ev_strides[j];
    }
    if (hi >= table_entries) return fail(PGX_ERR_BOUNDS, "gather range outside the table");
    if (n >= (1LL << 31)) return fail(PGX_ERR_UNSUPPORTED, "table too large");
    cudaStream_t st = (cudaStream_t)stream;
    // small descriptor arrays go to the device once per call
    int32_t h[6 * MAX_AXES];
    int32_t* d = nullptr;
    std::memset(h, 0, sizeof(h));
    for (int a = 0; a < n_free; ++a) {
        h[a] = free_dims[a];
        h[MAX_AXES + a] = free_strides[a];
    }
    for (int j = 0; j < n_ev; ++j) {
        h[2 * MAX_AXES + j] = ev_slots[j];
        h[3 * MAX_AXES + j] = ev_strides[j];
        h[4 * MAX_AXES + j] = ev_cards[j];
    }
    // the kernel tiles the batch by a power of two: derive the tile count from THAT width (ldb = 24 would otherwise
    // launch one tile of 16 sets and leave rows 16.. unwritten)
    const int bt_log2 = ilog2_floor(ldb < 32 ? ldb : 32);
    const int64_t bt = 1LL << bt_log2;
    const int64_t b_tiles = (B + bt - 1) / bt;
    if (b_tiles > 65535) return fail(PGX_ERR_UNSUPPORTED, "batch too large");
    PGX_CUDA(cudaMallocAsync((void**)&d, sizeof(h), st));
    cudaError_t ce = cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, st);
    if (ce == cudaSuccess) {
        const int per_block = 256 >> bt_log2;
        dim3 grid((unsigned)((n + per_block - 1) / per_block), (unsigned)b_tiles);
        if (dtype == PGX_F64)
            k_evidence_gather<double><<<grid, 256, 0, st>>>((const double*)table, n_free, d, d + MAX_AXES, n_ev,
                                                            d + 2 * MAX_AXES, d + 3 * MAX_AXES, d + 4 * MAX_AXES, ev_states,
                                                            ev_row_len, (double*)dst, n, B, ldb, bt_log2);
        else
            k_evidence_gather<float><<<grid, 256, 0, st>>>((const float*)table, n_free, d, d + MAX_AXES, n_ev,
                                                           d + 2 * MAX_AXES, d + 3 * MAX_AXES, d + 4 * MAX_AXES, ev_states,
                                                           ev_row_len, (float*)dst, n, B, ldb, bt_log2);
        ce = cudaGetLastError();
    }
    cudaFreeAsync(d, st);  // stream ordered: after the kernel on every path
    if (ce != cudaSuccess) return fail(PGX_ERR_CUDA, std::string("pgx_evidence_reduce: ") + cudaGetErrorString(ce));
    return PGX_OK;
}

int pgx_argmax_rows(int32_t dtype, const void* src, int64_t n, int64_t B, int32_t* out, void* stream) {
    if (!src || !out || n < 1 || B <= 0) return fail(PGX_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + 127) / 128);
    if (dtype == PGX_F64)
        k_argmax_rows<double><<<grid, 128, 0, st>>>((const double*)src, n, B, out);
    else
        k_argmax_rows<float><<<grid, 128, 0, st>>>((const float*)src, n, B, out);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

int pgx_normalize(int32_t dtype, const void* src, int64_t n, int64_t ldb, void* out, int64_t out_row_len, int64_t B,
                  void* stream) {
    if (!src || !out || n < 1 || B <= 0 || ldb < B || out_row_len < n) return fail(PGX_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((B + 127) / 128);
    if (dtype == PGX_F64)
        k_normalize<double><<<grid, 128, 0, st>>>((const double*)src, n, ldb, (double*)out, out_row_len, B);
    else
        k_normalize<float><<<grid, 128, 0, st>>>((const float*)src, n, ldb, (float*)out, out_row_len, B);
    PGX_CUDA(cudaGetLastError());
    return PGX_OK;
}

}  // extern "C"
