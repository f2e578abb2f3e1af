// pgx_step.cuh — the fused product + sum-out element function shared by every contraction kernel.
//
// One call computes ONE output entry `o` of ONE step for ONE evidence set `b`:
//
//     out[o, b] = REDUCE_s PROD_k operand_k[idx_k(o, s) + evoff_k(b), b]   ( / PROD_j divisor_j[idx_j(o), b] )
//
// which is the reference's product (pgmpy/factors/discrete/DiscreteFactor.py:769-777) + marginalize (:400-408)
// + reduce (:599-614) + divide (:838-863) collapsed into one pass that never materialises the joint table.
// Work tables are laid out [entry][ldb] with the evidence set fastest, so the 32 lanes of a warp (32
// consecutive evidence sets at the same `o`) always touch 32 consecutive elements whatever the strides are,
// and all index arithmetic below is warp-uniform.
//
// The step record layout is documented in pgmpy_b200/plan.py. This header is plain C++ on purpose: it has no
// warp intrinsics, so tests/hostsim compiles the very same function with g++ to check indexing on the CPU
// (test infrastructure only; the shipped library has no host execution path).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PGX_HD __device__ __forceinline__
#else
#define PGX_HD inline
#endif

namespace pgx {

constexpr int HEADER_WORDS = 16;
constexpr int STEP_FIXED = 12;
constexpr int OP_FIXED = 6;
constexpr int SEG_WORDS = 8;
constexpr int MAX_OPS = 16;
constexpr int MAX_AXES = 24;
constexpr uint32_t MAGIC = 0x50475831u;
constexpr int FLAG_MAX = 1;
constexpr int FLAG_DIV = 2;
constexpr int SEG_NORMALIZE = 1;

#if defined(__CUDACC__)
__host__ __device__ __forceinline__
#else
inline
#endif
int64_t ld_i64(const int32_t* w) { return (int64_t)(uint32_t)w[0] | ((int64_t)w[1] << 32); }

template <typename T>
PGX_HD T neg_inf() {
    return (T)(-1.0) / (T)0.0;
}

// rec      : step record (shared or global memory)
// cst      : packed batch-invariant tables
// ws       : workspace base; entry e of evidence set b lives at ws[e * ldb + b]
// ev_row   : evidence states of this evidence set (n_ev int32), may be null when the step has no evidence terms
// ev_card  : cardinality per evidence slot (states are clamped into range for memory safety)
template <typename T, int MAXK>
PGX_HD T contract_elem(const int32_t* __restrict__ rec, const T* __restrict__ cst, const T* __restrict__ ws,
                       const int32_t* __restrict__ ev_row, const int32_t* __restrict__ ev_card, int64_t ldb, int64_t b,
                       uint32_t o) {
    const int A = rec[0];
    const int S = rec[1];
    const int K = rec[2];
    const int flags = rec[3];
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = rec + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;

    const T* ptr[MAXK];      // operand pointer at (o, s = 0) for this evidence set
    int64_t unit[MAXK];      // elements per table entry: ldb for work tables, 1 for const tables
    int32_t eoff[MAXK];      // entry offset inside the table

#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        eoff[k] = 0;
        unit[k] = 1;
        ptr[k] = cst;
    }
    // mixed-radix digits of o, last axis fastest
    uint32_t rem = o;
    for (int a = A - 1; a >= 0; --a) {
        const uint32_t d = (uint32_t)odims[a];
        const uint32_t q = rem / d;
        const int32_t digit = (int32_t)(rem - q * d);
        rem = q;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < K) eoff[k] += digit * ops[k * opw + OP_FIXED + a];
    }
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        if (k < K) {
            const int32_t* op = ops + k * opw;
            const int kind = op[0] & 0xFF;
            const int n_ev = op[3];
            int32_t e = eoff[k];
            if (n_ev > 0) {
                const int32_t* pairs = rec + op[4];
                for (int j = 0; j < n_ev; ++j) {
                    const int slot = pairs[2 * j];
                    int32_t st = ev_row[slot];
                    const int32_t card = ev_card[slot];
                    st = st < 0 ? 0 : (st >= card ? card - 1 : st);
                    e += st * pairs[2 * j + 1];
                }
            }
            const int64_t base = ld_i64(op + 1) + e;
            if (kind == 1) {
                unit[k] = ldb;
                ptr[k] = ws + base * ldb + b;
            } else {
                unit[k] = 1;
                ptr[k] = cst + base;
            }
        }
    }
    int n_mul = K;
    if (flags & FLAG_DIV) {
        // divisors are the trailing operands
        while (n_mul > 0 && (ops[(n_mul - 1) * opw] & 0x100)) --n_mul;
    }

    const bool use_max = (flags & FLAG_MAX) != 0;
    T acc = use_max ? neg_inf<T>() : (T)0;
    if (S == 0) {
        T prod = (T)1;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k < n_mul) prod *= *ptr[k];
        acc = prod;
    } else {
        const int32_t inner = sdims[S - 1];
        int64_t istr[MAXK];
#pragma unroll
        for (int k = 0; k < MAXK; ++k) istr[k] = (k < n_mul) ? (int64_t)ops[k * opw + OP_FIXED + A + S - 1] * unit[k] : 0;
        const uint32_t n_outer = (uint32_t)(ld_i64(rec + 6) / inner);
        for (uint32_t u = 0; u < n_outer; ++u) {
            int64_t ooff[MAXK];
#pragma unroll
            for (int k = 0; k < MAXK; ++k) ooff[k] = 0;
            if (S > 1) {
                uint32_t r = u;
                for (int a = S - 2; a >= 0; --a) {
                    const uint32_t d = (uint32_t)sdims[a];
                    const uint32_t q = r / d;
                    const int32_t digit = (int32_t)(r - q * d);
                    r = q;
#pragma unroll
                    for (int k = 0; k < MAXK; ++k)
                        if (k < n_mul) ooff[k] += (int64_t)(digit * ops[k * opw + OP_FIXED + A + a]) * unit[k];
                }
            }
            for (int32_t j = 0; j < inner; ++j) {
                T prod = (T)1;
#pragma unroll
                for (int k = 0; k < MAXK; ++k)
                    if (k < n_mul) prod *= ptr[k][ooff[k] + (int64_t)j * istr[k]];
                if (use_max)
                    acc = prod > acc ? prod : acc;
                else
                    acc += prod;
            }
        }
    }
    if (flags & FLAG_DIV) {
        T den = (T)1;
#pragma unroll
        for (int k = 0; k < MAXK; ++k)
            if (k >= n_mul && k < K) den *= *ptr[k];
        T r = acc / den;
        acc = (r != r) ? (T)0 : r;  // 0/0 -> 0 ; x/0 stays inf (DiscreteFactor.py:859-863)
    }
    return acc;
}

// A run of consecutive output entries [o_begin, o_end) of one step for one evidence set, results stored to
// out[(o) * ldb] (out already points at this evidence set's column of the step's output table).
// Compared with calling contract_elem per entry, the mixed-radix decomposition of `o` is done only when the
// fastest output axis wraps (otherwise every operand offset advances by one stride), evidence offsets are
// resolved once, and the summed range walks its two fastest axes incrementally.
template <typename T, int MAXK>
PGX_HD void contract_run(const int32_t* __restrict__ rec, const T* __restrict__ cst, const T* __restrict__ ws,
                         const int32_t* __restrict__ ev_row, const int32_t* __restrict__ ev_card, int64_t ldb, int64_t b,
                         uint32_t o_begin, uint32_t o_end, T* __restrict__ out) {
    const int A = rec[0];
    const int S = rec[1];
    const int K = rec[2];
    const int flags = rec[3];
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = rec + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;

    const T* base[MAXK];   // table base for this evidence set (evidence offsets folded in)
    int64_t unit[MAXK];    // elements per table entry: ldb (work) or 1 (const)
    int32_t eoff[MAXK];    // entry offset of the current output entry
    int32_t ostep[MAXK];   // stride of the fastest output axis
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        base[k] = cst;
        unit[k] = 1;
        eoff[k] = 0;
        ostep[k] = 0;
        if (k < K) {
            const int32_t* op = ops + k * opw;
            int64_t e = ld_i64(op + 1);
            const int n_ev = op[3];
            if (n_ev > 0) {
                const int32_t* pairs = rec + op[4];
                for (int j = 0; j < n_ev; ++j) {
                    const int slot = pairs[2 * j];
                    int32_t st = ev_row[slot];
                    const int32_t card = ev_card[slot];
                    st = st < 0 ? 0 : (st >= card ? card - 1 : st);
                    e += st * pairs[2 * j + 1];
                }
            }
            if ((op[0] & 0xFF) == 1) {
                unit[k] = ldb;
                base[k] = ws + e * ldb + b;
            } else {
                base[k] = cst + e;
            }
            if (A > 0) ostep[k] = op[OP_FIXED + A - 1];
        }
    }
    int n_mul = K;
    if (flags & FLAG_DIV)
        while (n_mul > 0 && (ops[(n_mul - 1) * opw] & 0x100)) --n_mul;
    const bool use_max = (flags & FLAG_MAX) != 0;
    const uint32_t dim_last = A > 0 ? (uint32_t)odims[A - 1] : 1u;
    uint32_t d_last = dim_last;  // forces a full decomposition for the first entry

    // summed range: innermost axis is a plain loop, the next one advances incrementally
    const int32_t inner = S > 0 ? sdims[S - 1] : 1;
    const uint32_t n_outer = S > 0 ? (uint32_t)(ld_i64(rec + 6) / inner) : 1u;
    const uint32_t dim_s2 = S > 1 ? (uint32_t)sdims[S - 2] : 1u;
    int64_t istr[MAXK], s2str[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        istr[k] = (S > 0 && k < n_mul) ? (int64_t)ops[k * opw + OP_FIXED + A + S - 1] * unit[k] : 0;
        s2str[k] = (S > 1 && k < n_mul) ? (int64_t)ops[k * opw + OP_FIXED + A + S - 2] * unit[k] : 0;
    }

    for (uint32_t o = o_begin; o < o_end; ++o) {
        if (d_last >= dim_last) {
            uint32_t rem = o;
#pragma unroll
            for (int k = 0; k < MAXK; ++k) eoff[k] = 0;
            for (int a = A - 1; a >= 0; --a) {
                const uint32_t d = (uint32_t)odims[a];
                const uint32_t q = rem / d;
                const int32_t digit = (int32_t)(rem - q * d);
                if (a == A - 1) d_last = (uint32_t)digit;
                rem = q;
#pragma unroll
                for (int k = 0; k < MAXK; ++k)
                    if (k < K) eoff[k] += digit * ops[k * opw + OP_FIXED + a];
            }
            if (A == 0) d_last = 0;
        } else {
#pragma unroll
            for (int k = 0; k < MAXK; ++k) eoff[k] += ostep[k];
        }
        ++d_last;
        const T* ptr[MAXK];
#pragma unroll
        for (int k = 0; k < MAXK; ++k) ptr[k] = base[k] + (int64_t)eoff[k] * unit[k];

        T acc = use_max ? neg_inf<T>() : (T)0;
        if (S == 0) {
            T prod = (T)1;
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
                if (k < n_mul) prod *= *ptr[k];
            acc = prod;
        } else {
            int64_t ooff[MAXK];
#pragma unroll
            for (int k = 0; k < MAXK; ++k) ooff[k] = 0;
            uint32_t d2 = dim_s2;  // full decomposition on the first outer iteration
            for (uint32_t u = 0; u < n_outer; ++u) {
                if (S > 1) {
                    if (d2 >= dim_s2) {
                        uint32_t r = u;
#pragma unroll
                        for (int k = 0; k < MAXK; ++k) ooff[k] = 0;
                        for (int a = S - 2; a >= 0; --a) {
                            const uint32_t d = (uint32_t)sdims[a];
                            const uint32_t q = r / d;
                            const int32_t digit = (int32_t)(r - q * d);
                            if (a == S - 2) d2 = (uint32_t)digit;
                            r = q;
#pragma unroll
                            for (int k = 0; k < MAXK; ++k)
                                if (k < n_mul) ooff[k] += (int64_t)(digit * ops[k * opw + OP_FIXED + A + a]) * unit[k];
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < MAXK; ++k) ooff[k] += s2str[k];
                    }
                    ++d2;
                }
                for (int32_t j = 0; j < inner; ++j) {
                    T prod = (T)1;
#pragma unroll
                    for (int k = 0; k < MAXK; ++k)
                        if (k < n_mul) prod *= ptr[k][ooff[k] + (int64_t)j * istr[k]];
                    if (use_max)
                        acc = prod > acc ? prod : acc;
                    else
                        acc += prod;
                }
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

// Dispatch on the operand count so narrow steps do not pay for MAX_OPS-wide unrolled loops.
// LIMIT caps the widest instantiation compiled into the calling kernel (register budget).
template <typename T, int LIMIT>
PGX_HD T contract_elem_upto(const int32_t* __restrict__ rec, const T* __restrict__ cst, const T* __restrict__ ws,
                            const int32_t* __restrict__ ev_row, const int32_t* __restrict__ ev_card, int64_t ldb,
                            int64_t b, uint32_t o) {
    const int K = rec[2];
    if (LIMIT <= 2 || K <= 2) return contract_elem<T, 2>(rec, cst, ws, ev_row, ev_card, ldb, b, o);
    if (LIMIT <= 4 || K <= 4) return contract_elem<T, (LIMIT < 4 ? LIMIT : 4)>(rec, cst, ws, ev_row, ev_card, ldb, b, o);
    if (LIMIT <= 8 || K <= 8) return contract_elem<T, (LIMIT < 8 ? LIMIT : 8)>(rec, cst, ws, ev_row, ev_card, ldb, b, o);
    return contract_elem<T, (LIMIT < MAX_OPS ? LIMIT : MAX_OPS)>(rec, cst, ws, ev_row, ev_card, ldb, b, o);
}

}  // namespace pgx
