// pgx_stage.cuh — K3: GEMM-shaped contraction steps, operands staged through shared memory by the TMA engine.
//
//     out[o, b] = sum_s  P[ip(o, s), b] * Q[iq(o, s), b]                (two operands, plain sum-product)
//
// where each operand depends on only PART of the output scope (diabetes: out (ins_sens, cho_4, bg_4) = sum over
// (cho_bal_2, gut_abs_3) of a message over (cho_bal_2, gut_abs_3, cho_4) times a message over (ins_sens, cho_bal_2,
// gut_abs_3, bg_4): a 21 x 55 x K96 matrix product per evidence set). The streaming kernel (k_contract_tile32) issues
// one 256-byte operand row load per multiply; its throughput is the L1 data path (128 B/clk/SM), 5-16 % of the fp64
// pipe. Here the reuse a GEMM tile gets is made explicit:
//
//   * the host picks two output axes X, Y such that one operand does not depend on Y and the other not on X (or at least
//     one of them does not depend on one axis), a CTA owns a box tile of (4 bx) x (4 by) outputs over (X, Y) — the other
//     output axes are fixed per CTA — for ONE tile of 32 evidence sets (lane = evidence set, as everywhere);
//   * the rows of each operand that the tile touches are copied ONCE per CTA into shared memory by bulk asynchronous
//     copies (cp.async.bulk.shared.global with mbarrier complete_tx — the 1-D TMA path: UBLKCP + SYNCS in SASS), the
//     summed range cut into chunks of `sc` indices, two stages in flight (copy of chunk c + 1 overlaps the math of c);
//   * a warp owns one 4 x 4 register block of outputs: per summed index it reads 4 + 4 rows from shared memory for 16
//     fused multiply-adds (an X-only times a Y-only operand), instead of 32 global loads.
//
// Batch-invariant operands (CPTs) have no evidence-set dimension: they are staged as scalars and read with uniform
// (broadcast) shared-memory loads through the same code path (row pitch 1, lane multiplier 0).
//
// Operand dependence forms, after the host's choice of axes and operand order (TP for operand P, TQ for Q;
// 0 = neither axis, 1 = X only, 2 = Y only, 3 = both):  (1,2) GEMM | (1,3) | (0,3) | (0,1) | (1,1).
#pragma once
#include <cstdint>

#include "pgx_step.cuh"

namespace pgx {

// One (step, tiling) entry of a staged launch.
struct StageItem {
    int32_t rec_off, rec_len;
    int32_t ax, ay;     // output axes of the tile (ay = -1: one-dimensional tile)
    int32_t bx, by;     // register blocks (4 x 4 outputs) per tile along X and Y, bx * by <= 8
    int32_t ntx, nty;   // tiles along X and Y
    int32_t tiles;      // CTAs per tile of 32 evidence sets = (product of the other axes) * ntx * nty
    int32_t sc;         // summed indices per pipeline stage
    int32_t swap;       // 1: operand 1 of the step record plays P, operand 0 plays Q
    int32_t form;       // TP * 4 + TQ
    int32_t blk_begin;  // first linear CTA index of this step inside the launch
    int32_t stage_elems;  // elements (of T) per pipeline stage
    int32_t p_elems;      // elements of the P region inside a stage (Q follows)
    int32_t pad;
};

#if defined(__CUDACC__)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk asynchronous copy global -> shared, completion counted in bytes on the mbarrier (TMA engine, no tensor map)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

struct StageOp {
    uint32_t gbase;   // element index of the operand at the tile origin, summed index 0 (u32 address space of ws_all)
    uint32_t gx, gy;  // element stride per step along X / Y (0: independent)
    uint32_t gunit;   // elements per table entry: ldb (work table) or 1 (batch-invariant table)
    int32_t nx, ny;   // staged extents along X / Y (tile extent where the operand depends on the axis, else 1)
    int32_t pitch;    // elements per staged slot: 32 (work: one row of evidence sets) or 1 (batch-invariant scalar)
    int32_t is_const;
};

// TP / TQ: dependence of operand P / Q on the tile axes (bit 0 = X, bit 1 = Y).
//   sp / sq : the operand's region of the current stage (lane already folded in for work tables)
//   slot(x, y, s) = (x * ny + y) * sc + s ; element = slot * pitch
template <typename T, int TP, int TQ>
__device__ __forceinline__ void stage_compute(const T* __restrict__ sp, const T* __restrict__ sq, int sc, int n_s, int pny,
                                              int qny, int p_pitch, int q_pitch, int xl0, int yl0, T (&acc)[4][4]) {
    constexpr bool PX = (TP & 1) != 0, PY = (TP & 2) != 0, QX = (TQ & 1) != 0, QY = (TQ & 2) != 0;
    int po[PX ? 4 : 1][PY ? 4 : 1], qo[QX ? 4 : 1][QY ? 4 : 1];
#pragma unroll
    for (int i = 0; i < (PX ? 4 : 1); ++i)
#pragma unroll
        for (int j = 0; j < (PY ? 4 : 1); ++j) po[i][j] = (((PX ? xl0 + i : 0) * pny) + (PY ? yl0 + j : 0)) * sc * p_pitch;
#pragma unroll
    for (int i = 0; i < (QX ? 4 : 1); ++i)
#pragma unroll
        for (int j = 0; j < (QY ? 4 : 1); ++j) qo[i][j] = (((QX ? xl0 + i : 0) * qny) + (QY ? yl0 + j : 0)) * sc * q_pitch;
#pragma unroll 2
    for (int s = 0; s < n_s; ++s) {
        T pv[PX ? 4 : 1][PY ? 4 : 1], qv[QX ? 4 : 1][QY ? 4 : 1];
#pragma unroll
        for (int i = 0; i < (PX ? 4 : 1); ++i)
#pragma unroll
            for (int j = 0; j < (PY ? 4 : 1); ++j) pv[i][j] = sp[po[i][j] + s * p_pitch];
#pragma unroll
        for (int i = 0; i < (QX ? 4 : 1); ++i)
#pragma unroll
            for (int j = 0; j < (QY ? 4 : 1); ++j) qv[i][j] = sq[qo[i][j] + s * q_pitch];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                acc[i][j] = fma(pv[PX ? i : 0][PY ? j : 0], qv[QX ? i : 0][QY ? j : 0], acc[i][j]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 2) k_contract_stage(const int32_t* __restrict__ pool,
                                                           const StageItem* __restrict__ items, int n_items,
                                                           const T* __restrict__ ws_in, T* __restrict__ ws_out,
                                                           uint32_t ws_off0, int64_t B, uint32_t ldb) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (items[mid].blk_begin <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const StageItem it = items[lo];
    const int local = (int)blockIdx.x - it.blk_begin;
    const int b_tile = local / it.tiles;
    int tile = local - b_tile * it.tiles;
    const int ty_i = tile % it.nty;
    tile /= it.nty;
    const int tx_i = tile % it.ntx;
    const int other = tile / it.ntx;

    // shared memory: [2 mbarriers][StageOp x 2][out descriptor][step record][stab: S x 2][stage 0][stage 1]
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_raw);
    StageOp* sops = reinterpret_cast<StageOp*>(s_raw + 16);
    uint32_t* s_out = reinterpret_cast<uint32_t*>(s_raw + 16 + 2 * sizeof(StageOp));  // [0] base, [1] ox, [2] oy, [3] ext x, [4] ext y
    int32_t* s_rec = reinterpret_cast<int32_t*>(s_raw + 128);
    for (int i = threadIdx.x; i < it.rec_len; i += blockDim.x) s_rec[i] = pool[it.rec_off + i];
    __syncthreads();
    const int A = s_rec[0], S = s_rec[1];
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = s_rec + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;
    const int sum_size = s_rec[6];
    int32_t* s_stab = s_rec + ((it.rec_len + 3) & ~3);
    T* stage0 = reinterpret_cast<T*>(s_raw + 128 + (size_t)(((it.rec_len + 3) & ~3) + ((2 * sum_size + 3) & ~3)) * 4);
    // the stages must start on a 16-byte boundary (bulk copies): 128 + multiples of 16 bytes above
    const uint32_t b0 = (uint32_t)b_tile * 32u;

    if (threadIdx.x == 0) {
        // tile origin: digits of `other` over the axes that are not tile axes (last axis fastest)
        uint32_t rem = (uint32_t)other;
        int32_t fix[2] = {0, 0};
        uint32_t fixo = 0, ostride = 1, ox = 0, oy = 0;
        for (int a = A - 1; a >= 0; --a) {
            const uint32_t d = (uint32_t)odims[a];
            if (a == it.ax) ox = ostride;
            else if (a == it.ay) oy = ostride;
            else {
                const uint32_t q = rem / d;
                const uint32_t digit = rem - q * d;
                rem = q;
                fixo += digit * ostride;
                for (int k = 0; k < 2; ++k) fix[k] += (int32_t)digit * ops[k * opw + OP_FIXED + a];
            }
            ostride *= d;
        }
        const int x0 = tx_i * 4 * it.bx, y0 = ty_i * 4 * it.by;
        const int ext_x = min(4 * it.bx, odims[it.ax] - x0);
        const int ext_y = it.ay >= 0 ? min(4 * it.by, odims[it.ay] - y0) : 1;
        for (int r = 0; r < 2; ++r) {  // r = 0: P, r = 1: Q
            const int k = r ^ it.swap;
            const int32_t* op = ops + k * opw;
            const int32_t sx = op[OP_FIXED + it.ax], sy = it.ay >= 0 ? op[OP_FIXED + it.ay] : 0;
            const bool work = (op[0] & 0xFF) == 1;
            StageOp so;
            so.gunit = work ? ldb : 1u;
            const uint32_t e = (uint32_t)op[1] + (uint32_t)fix[k] + (uint32_t)(x0 * sx) + (uint32_t)(y0 * sy);
            so.gbase = work ? ws_off0 + e * ldb + b0 : e;
            so.gx = (uint32_t)sx * so.gunit;
            so.gy = (uint32_t)sy * so.gunit;
            so.nx = sx ? ext_x : 1;
            so.ny = sy ? ext_y : 1;
            so.pitch = work ? 32 : 1;
            so.is_const = work ? 0 : 1;
            sops[r] = so;
        }
        s_out[0] = ws_off0 + ((uint32_t)s_rec[8] + fixo + (uint32_t)x0 * ox + (uint32_t)y0 * oy) * ldb + b0;
        s_out[1] = ox * ldb;
        s_out[2] = oy * ldb;
        s_out[3] = (uint32_t)ext_x;
        s_out[4] = (uint32_t)ext_y;
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // summed-range offset table, in table entries: stab[q][r]
    for (int qi = threadIdx.x; qi < sum_size; qi += blockDim.x) {
        uint32_t rem = (uint32_t)qi;
        int32_t off[2] = {0, 0};
        for (int a = S - 1; a >= 0; --a) {
            const uint32_t d = (uint32_t)sdims[a];
            const uint32_t q = rem / d;
            const int32_t digit = (int32_t)(rem - q * d);
            rem = q;
            off[0] += digit * ops[(0 ^ it.swap) * opw + OP_FIXED + A + a];
            off[1] += digit * ops[(1 ^ it.swap) * opw + OP_FIXED + A + a];
        }
        s_stab[2 * qi] = off[0];
        s_stab[2 * qi + 1] = off[1];
    }
    __syncthreads();
    const StageOp P = sops[0], Q = sops[1];
    const int sc = it.sc;
    const int n_chunks = (sum_size + sc - 1) / sc;
    // the staged layout uses the NOMINAL tile extent 4 by along Y so that slot arithmetic is the same in every tile
    const int pny = P.gy ? 4 * it.by : 1;
    const int qny = Q.gy ? 4 * it.by : 1;

    auto issue = [&](int c) {
        T* st = stage0 + (size_t)(c & 1) * it.stage_elems;
        const int q0 = c * sc;
        const int n_s = min(sc, sum_size - q0);
        uint64_t* bar = &bars[c & 1];
        if (it.pad & 2) return;  // (tuning aid: no copies)
        if (threadIdx.x == 0) {
            uint32_t bytes = 0;
            if (!P.is_const) bytes += (uint32_t)(P.nx * P.ny * n_s) * 32u * (uint32_t)sizeof(T);
            if (!Q.is_const) bytes += (uint32_t)(Q.nx * Q.ny * n_s) * 32u * (uint32_t)sizeof(T);
            mbar_expect_tx(bar, bytes);
        }
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const StageOp& O = r ? Q : P;
            T* reg = st + (r ? it.p_elems : 0);
            const int ny_nom = r ? qny : pny;
            const int n = O.nx * O.ny * n_s;
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                const int si = j % n_s;
                const int t = j / n_s;
                const int yi = t % O.ny, xi = t / O.ny;
                const uint32_t g = O.gbase + (uint32_t)xi * O.gx + (uint32_t)yi * O.gy +
                                   (uint32_t)s_stab[2 * (q0 + si) + r] * O.gunit;
                const int slot = (xi * ny_nom + yi) * sc + si;
                if (O.is_const)
                    reg[slot] = ws_in[g];
                else
                    bulk_g2s(reg + (size_t)slot * 32, ws_in + g, 32u * (uint32_t)sizeof(T), bar);
            }
        }
    };

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bxi = warp / it.by, byi = warp - bxi * it.by;
    const bool active = bxi < it.bx;
    const int xl0 = 4 * bxi, yl0 = 4 * byi;
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = (T)0;

    issue(0);
    for (int c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) issue(c + 1);
        if (!(it.pad & 2)) mbar_wait(&bars[c & 1], (uint32_t)((c >> 1) & 1));
        __syncthreads();  // batch-invariant scalars are written with ordinary stores
        if (active && !(it.pad & 1)) {
            const T* st = stage0 + (size_t)(c & 1) * it.stage_elems;
            const T* sp = st + (P.is_const ? 0 : lane);
            const T* sq = st + it.p_elems + (Q.is_const ? 0 : lane);
            const int n_s = min(sc, sum_size - c * sc);
            switch (it.form) {
                case 1 * 4 + 2: stage_compute<T, 1, 2>(sp, sq, sc, n_s, pny, qny, P.pitch, Q.pitch, xl0, yl0, acc); break;
                case 1 * 4 + 3: stage_compute<T, 1, 3>(sp, sq, sc, n_s, pny, qny, P.pitch, Q.pitch, xl0, yl0, acc); break;
                case 0 * 4 + 3: stage_compute<T, 0, 3>(sp, sq, sc, n_s, pny, qny, P.pitch, Q.pitch, xl0, yl0, acc); break;
                case 0 * 4 + 1: stage_compute<T, 0, 1>(sp, sq, sc, n_s, pny, qny, P.pitch, Q.pitch, xl0, yl0, acc); break;
                default: stage_compute<T, 1, 1>(sp, sq, sc, n_s, pny, qny, P.pitch, Q.pitch, xl0, yl0, acc); break;
            }
        }
        __syncthreads();  // every warp is done with this stage before the copy of chunk c + 2 may overwrite it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (active && (int64_t)b0 + lane < B) {
        const uint32_t ob = s_out[0] + (uint32_t)lane, ox = s_out[1], oy = s_out[2];
        const int ext_x = (int)s_out[3], ext_y = (int)s_out[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (xl0 + i < ext_x && yl0 + j < ext_y) ws_out[ob + (uint32_t)(xl0 + i) * ox + (uint32_t)(yl0 + j) * oy] = acc[i][j];
    }
}

#endif  // __CUDACC__

}  // namespace pgx
