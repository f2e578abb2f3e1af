// pgx_mm.cu — K3: matrix-product-shaped contraction steps (reference: the einsum pair
// pgmpy/factors/discrete/DiscreteFactor.py:769-777 (product) + :400-408 (sum-out) of two factors sharing the summed
// variables), sm_100a.
//
//     out[o, b] = sum_s P[ip(o, s), b] * Q[iq(o, s), b]          b = evidence set, fastest in memory ([entry][ldb])
//
// Per evidence set this is a batch of Z small matrix products out_z[M, N] = P_z[M, K] Q_z[K, N] (diabetes: 55 x 90 x K17,
// 96 x 55 x K21 ...; munin: 500 products of 16 x 49 x K7). Lane = evidence set, so a "row" is 32 consecutive evidence
// sets of one table entry (256 bytes in fp64) and every address is warp uniform. The streaming step kernel
// (k_contract_tile32) issues one row load per multiply and ran these steps at 10-25 % of the HBM roofline, bound by L1
// wavefronts and load latency. Here the reuse of a matrix product is made explicit:
//
//   * a CTA walks a run of consecutive tiles of TZ x TX x TY outputs for one tile of 32 evidence sets;
//   * the operand rows of (tile, chunk of KC summed indices) go into a ring of shared-memory stages by asynchronous
//     copies (cp.async 16 B -> LDGSTS, completion counted on the stage's mbarrier: SYNCS / ARRIVES.LDGSTSBAR in SASS),
//     issued n_stages - 2 chunks ahead of the math and across tile boundaries, every thread copying its share (a
//     dedicated producer warp would leave 96 registers per thread: the 17th warp lands on one scheduler, and
//     setmaxnreg cannot fix that because ptxas does not confine the producer code to the reduced count); row
//     addresses come from offset tables the host tabulated per step;
//   * the 8 warps of a CTA (two CTAs per SM) each own a 4 x 8 register block of the tile: per summed index 4 + 8 shared-memory row reads
//     feed 32 fused multiply-adds (the streaming kernel: 64 loads), full/empty mbarriers per stage, no __syncthreads
//     in the pipeline;
//   * a batch-invariant P (a CPT: "[M x K] . [K x (N . B)]", the evidence batch folded into the matrix N dimension) is
//     never staged: in fp64 the consumer warps run it on the tensor cores — mma.sync m8n8k4 f64 (DMMA), A fragments from
//     the table copy, B fragments = 4 summed indices x 8 consecutive evidence sets straight out of the staged rows — and
//     in fp32 with broadcast loads + FFMA.
#include "pgx_mm.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pgx_step.cuh"

namespace pgx {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
// 16-byte asynchronous copy global (L2) -> shared, no register staging (LDGSTS). No "memory" clobber on purpose: the
// copies are ordered against their consumers by the mbarrier operations (which do clobber), and a clobber here would
// force every loop-invariant value to be re-read between two copies.
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src));
}
// deferred arrival: fires on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// Everything of an MMItem the pipeline loop touches, unpacked into registers once per CTA.
struct MMCfg {
    int lgTX, lgTY, lgKC, TZ;
    int M, N, Z, K;
    int n_tiles, ntx, nty, n_chunks, n_stages;
    int p_rows1, rows1;  // staged rows per summed index: P rows (0 when P is batch invariant), P + Q rows
    int slab;            // elements between consecutive summed indices inside a stage
    int stage_elems;
    uint32_t p_base, q_base;
};

struct TileCoord {
    int x0, y0, z0;
    uint32_t b0;
};

__device__ __forceinline__ TileCoord tile_coord(const MMCfg& cf, int64_t t) {
    const int bt = (int)(t / cf.n_tiles);
    int r = (int)(t - (int64_t)bt * cf.n_tiles);
    const int ty = r % cf.nty;
    r /= cf.nty;
    const int tx = r % cf.ntx;
    const int tz = r / cf.ntx;
    TileCoord c;
    c.x0 = tx << cf.lgTX;
    c.y0 = ty << cf.lgTY;
    c.z0 = tz * cf.TZ;
    c.b0 = (uint32_t)bt * 32u;
    return c;
}

// Stage layout: [summed index sl][row][32 evidence sets]; row = the tile's P rows (zl * TX + x) followed by its Q rows
// (zl * TY + y), so one formula addresses both operands: element (sl * slab + row * 32 + b). A tensor-core step skews
// consecutive summed indices by 8 elements (slab = rows1 * 32 + 8) so that a DMMA B fragment (4 summed indices x 8
// evidence sets) reads 32 distinct 8-byte words in two wavefronts.
constexpr int MMA_SKEW = 8;

// Producer side: every thread of the CTA copies its share of each stage with 16-byte asynchronous copies (cp.async ->
// LDGSTS, L2 -> shared memory without passing registers); a row of 32 evidence sets is 16 (fp64) or 8 (fp32) such
// pieces. Completion is counted on the stage's `full` mbarrier by cp.async.mbarrier.arrive.noinc (one deferred arrival
// per thread and stage), so nobody blocks on its own copies. The first version used one bulk copy (cp.async.bulk, the
// TMA engine) per 256-byte row: measured 41 cycles per request and SM — 1.8 TB/s for the whole chip — which made the
// copies 5x more expensive than the math (profiles/r02_mm_kernel.md); LDGSTS moves 512 bytes per warp instruction.
// Thread t always serves piece (t mod pieces-per-row) of rows (t / pieces-per-row) + j * rows-per-pass: the source
// row of assignment j at summed index 0 is computed once per tile (src[j]); a chunk adds only the warp-uniform
// summed-index offsets, and the shared-memory address of assignment j is the thread's first one plus j * 8 KB.
constexpr int MM_MAXJ = 6;  // assignments per thread and summed index: rows per summed index <= 192, >= 32 rows per pass

template <typename T>
struct Feeder {
    static constexpr int LG_PPR = sizeof(T) == 8 ? 4 : 3;   // log2(16-byte pieces per row)
    static constexpr int EPP = 16 / (int)sizeof(T);         // elements per piece
    static constexpr int ROWS_PER_PASS = MM_THREADS >> LG_PPR;
    const MMCfg& cf;
    const int32_t *xoffP, *yoffQ, *soffP, *soffQ, *zoffP, *zoffQ;  // soffP / soffQ already multiplied by ldb on the host
    const T* ws_in;
    T* stages;
    uint64_t *full, *empty;
    uint32_t ws_off0, ldb;
    int no_copies;  // tuning aid
    int64_t t;      // tile of the next chunk to issue
    int c;          // chunk inside that tile
    int stage;
    uint32_t phase;
    uint32_t src[MM_MAXJ];

    __device__ __forceinline__ void load_tile() {
        const TileCoord tc = tile_coord(cf, t);
        const int tid = threadIdx.x;
        const uint32_t col = ws_off0 + tc.b0 + (uint32_t)(tid & ((1 << LG_PPR) - 1)) * EPP;
#pragma unroll
        for (int j = 0; j < MM_MAXJ; ++j) {
            const int row = (tid >> LG_PPR) + j * ROWS_PER_PASS;
            src[j] = 0;
            if (row < cf.p_rows1) {
                const int x = tc.x0 + (row & ((1 << cf.lgTX) - 1)), zl = tc.z0 + (row >> cf.lgTX);
                src[j] = col + (cf.p_base + (uint32_t)(__ldg(zoffP + (zl < cf.Z ? zl : cf.Z - 1)) + __ldg(xoffP + (x < cf.M ? x : cf.M - 1)))) * ldb;
            } else if (row < cf.rows1) {
                const int r2 = row - cf.p_rows1;
                const int y = tc.y0 + (r2 & ((1 << cf.lgTY) - 1)), zl = tc.z0 + (r2 >> cf.lgTY);
                src[j] = col + (cf.q_base + (uint32_t)(__ldg(zoffQ + (zl < cf.Z ? zl : cf.Z - 1)) + __ldg(yoffQ + (y < cf.N ? y : cf.N - 1)))) * ldb;
            } else {
                break;
            }
        }
    }

    __device__ __forceinline__ void issue() {
        const int tid = threadIdx.x;
        mbar_wait(&empty[stage], phase ^ 1u);
        if (!no_copies) {
            const int k0 = c << cf.lgKC;
            const int n_s = cf.K - k0 < (1 << cf.lgKC) ? cf.K - k0 : (1 << cf.lgKC);
            const int row0 = tid >> LG_PPR;
            const uint32_t dst0 = smem_u32(stages + (size_t)stage * cf.stage_elems) + (uint32_t)row0 * (32u * (uint32_t)sizeof(T)) +
                                  (uint32_t)(tid & ((1 << LG_PPR) - 1)) * 16u;
            const uint32_t step = (uint32_t)cf.slab * (uint32_t)sizeof(T);
            for (int sl = 0; sl < n_s; ++sl) {
                const uint32_t sop = (uint32_t)__ldg(soffP + k0 + sl), soq = (uint32_t)__ldg(soffQ + k0 + sl);
                const uint32_t d = dst0 + (uint32_t)sl * step;
#pragma unroll
                for (int j = 0; j < MM_MAXJ; ++j) {
                    const int row = row0 + j * ROWS_PER_PASS;
                    if (row >= cf.rows1) break;
                    cp_async16(d + (uint32_t)j * (ROWS_PER_PASS * 32u * (uint32_t)sizeof(T)), ws_in + src[j] + (row < cf.p_rows1 ? sop : soq));
                }
            }
        }
        cp_async_arrive_noinc(&full[stage]);  // arrives when this thread's copies above have landed
        if (++stage == cf.n_stages) {
            stage = 0;
            phase ^= 1u;
        }
        if (++c == cf.n_chunks) {
            c = 0;
            ++t;
            load_tile();
        }
    }
};

template <typename T, int KC, bool PCONST>
__device__ __forceinline__ void fma_chunk(const T* __restrict__ sp, const T* __restrict__ sq, int slab,
                                          const T* __restrict__ cst, const uint32_t (&pidx)[4],
                                          const int32_t* __restrict__ soffP, int n_s, T (&acc)[4][8]) {
    if (n_s == KC) {
#pragma unroll
        for (int sl = 0; sl < KC; ++sl) {
            T p[4], q[8];
            if (PCONST) {
                const uint32_t so = (uint32_t)__ldg(soffP + sl);
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = __ldg(cst + pidx[i] + so);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = sp[sl * slab + i * 32];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = sq[sl * slab + j * 32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(p[i], q[j], acc[i][j]);
        }
    } else {
#pragma unroll 1
        for (int sl = 0; sl < n_s; ++sl) {
            T p[4], q[8];
            if (PCONST) {
                const uint32_t so = (uint32_t)__ldg(soffP + sl);
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = __ldg(cst + pidx[i] + so);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) p[i] = sp[sl * slab + i * 32];
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) q[j] = sq[sl * slab + j * 32];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(p[i], q[j], acc[i][j]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(MM_THREADS, MM_CTAS_PER_SM) k_contract_mm(const MMItem* __restrict__ items, int n_items,
                                                                const int32_t* __restrict__ tabs,
                                                                const T* __restrict__ ws_in, T* __restrict__ ws_out,
                                                                uint32_t ws_off0, int64_t B, uint32_t ldb, int b_tiles) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&items[mid].blk_begin) <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const MMItem* __restrict__ ip = items + lo;
    MMCfg cf;
    cf.lgTX = ip->lgTX, cf.lgTY = ip->lgTY, cf.lgKC = ip->lgKC, cf.TZ = ip->TZ;
    cf.M = ip->M, cf.N = ip->N, cf.Z = ip->Z, cf.K = ip->K;
    cf.n_tiles = ip->n_tiles, cf.ntx = ip->ntx, cf.nty = ip->nty, cf.n_chunks = ip->n_chunks, cf.n_stages = ip->n_stages;
    const int p_const = ip->p_const, use_mma = ip->use_mma, n_active = ip->n_active, dbg = ip->pad0;
    cf.p_rows1 = p_const ? 0 : (cf.TZ << cf.lgTX);
    cf.rows1 = cf.p_rows1 + (cf.TZ << cf.lgTY);
    cf.slab = cf.rows1 * 32 + (use_mma ? MMA_SKEW : 0);
    cf.stage_elems = ip->stage_elems;
    cf.p_base = ip->p_base, cf.q_base = ip->q_base;
    const uint32_t o_base = ip->o_base;
    uint64_t* full = reinterpret_cast<uint64_t*>(s_raw);
    uint64_t* empty = full + MM_MAX_STAGES;
    T* stages = reinterpret_cast<T*>(s_raw + 128);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < cf.n_stages; ++s) {
            mbar_init(&full[s], MM_THREADS);
            mbar_init(&empty[s], (uint32_t)n_active);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t total = (int64_t)cf.n_tiles * b_tiles;
    const int64_t t0 = (int64_t)((int)blockIdx.x - ip->blk_begin) * ip->tiles_per_cta;
    const int64_t t1 = t0 + ip->tiles_per_cta < total ? t0 + ip->tiles_per_cta : total;
    const int TX = 1 << cf.lgTX, TY = 1 << cf.lgTY, KC = 1 << cf.lgKC;
    const int M = cf.M, N = cf.N, Z = cf.Z, K = cf.K;
    const int32_t* xoffP = tabs + ip->tab;
    const int32_t* xoffO = xoffP + M;
    const int32_t* yoffQ = xoffO + M;
    const int32_t* yoffO = yoffQ + N;
    const int32_t* soffP = yoffO + N;  // summed-index offsets in table entries ...
    const int32_t* soffQ = soffP + K;
    const int32_t* zoffP = soffQ + K;
    const int32_t* zoffQ = zoffP + Z;
    const int32_t* zoffO = zoffQ + Z;
    const int32_t* soffPl = zoffO + Z;  // ... and the same multiplied by ldb (elements), for the copies
    const int32_t* soffQl = soffPl + K;

    // the CTA's run as one sequence of chunks g = 0 .. G-1; chunk g + D is handed to the copy engine before chunk g is
    // consumed. D = n_stages - 2 leaves one chunk of slack between the warps (the refill of a stage then waits for a
    // chunk everybody finished one iteration ago); shallower rings refill the stage just released.
    const int64_t G = (t1 - t0) * cf.n_chunks;
    const int D = cf.n_stages >= 4 ? cf.n_stages - 2 : cf.n_stages - 1;
    Feeder<T> feed{cf, xoffP, yoffQ, p_const ? soffQl : soffPl, soffQl, zoffP, zoffQ, ws_in, stages, full, empty, ws_off0, ldb,
                   dbg & 2, t0, 0, 0, 0u, {}};
    feed.load_tile();
    for (int i = 0; i < D && i < G; ++i) feed.issue();
    const bool active = warp < n_active;
    int stage = 0;
    uint32_t phase = 0;
    int64_t g = 0;

    if (use_mma) {
        // fp64 tensor cores: block = 8 x (one DMMA M tile) x 4 y x 32 evidence sets (4 DMMA N tiles of 8 sets each).
        // A = P[x, k] (batch invariant, 8 x 4 per DMMA), B = Q[k][y, b] (4 x 8), C/D = out[x][y, b] (8 x 8).
        if constexpr (sizeof(T) == 8) {
            const int bxn = TX >> 3, byn = TY >> 2;
            const int blocks = bxn * byn;
            const int bz = warp / blocks, rem = warp - bz * blocks, bx = rem / byn, by = rem - bx * byn;
            const int gq = lane >> 2, tq = lane & 3;
            for (int64_t t = t0; t < t1; ++t) {
                const TileCoord tc = tile_coord(cf, t);
                const int z = tc.z0 + bz, zc = z < Z ? z : Z - 1;
                const int xa = tc.x0 + 8 * bx + gq, xac = xa < M ? xa : M - 1;
                const uint32_t aidx = cf.p_base + (uint32_t)(__ldg(zoffP + zc) + __ldg(xoffP + xac));
                double acc[4][4][2];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int n = 0; n < 4; ++n) acc[j][n][0] = acc[j][n][1] = 0.0;
                for (int c = 0; c < cf.n_chunks; ++c, ++g) {
                    if (g + D < G) feed.issue();
                    if (!active) continue;
                    mbar_wait(&full[stage], phase);
                    const int k0 = c << cf.lgKC;
                    const int n_s = K - k0 < KC ? K - k0 : KC;
                    const double* sq = reinterpret_cast<const double*>(stages) + (size_t)stage * cf.stage_elems +
                                       ((bz * TY + 4 * by) << 5) + gq;
                    if (!(dbg & 1)) {
                        for (int ks = 0; ks < n_s; ks += 4) {
                            const int sl = ks + tq;
                            const bool kin = sl < n_s;
                            const double a = kin ? __ldg(reinterpret_cast<const double*>(ws_in) + aidx + (uint32_t)__ldg(soffP + k0 + sl)) : 0.0;
                            const double* row = sq + (size_t)(kin ? sl : 0) * cf.slab;
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int n = 0; n < 4; ++n) {
                                    const double bv = kin ? row[(j << 5) + 8 * n] : 0.0;
                                    dmma884(acc[j][n][0], acc[j][n][1], a, bv);
                                }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
                    if (++stage == cf.n_stages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (active && z < Z && xa < M) {
                    const uint32_t zx = o_base + (uint32_t)(__ldg(zoffO + z) + __ldg(xoffO + xa));
                    double* ob = reinterpret_cast<double*>(ws_out) + ws_off0 + tc.b0 + 2 * tq;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int y = tc.y0 + 4 * by + j;
                        if (y < N) {
                            double* orow = ob + (size_t)(zx + (uint32_t)__ldg(yoffO + y)) * ldb;
#pragma unroll
                            for (int n = 0; n < 4; ++n)
                                *reinterpret_cast<double2*>(orow + 8 * n) = make_double2(acc[j][n][0], acc[j][n][1]);
                        }
                    }
                }
            }
        }
        return;
    }
    const int bxn = TX >> 2, byn = TY >> 3;
    const int blocks = bxn * byn;
    const int bz = warp / blocks, rem = warp - bz * blocks, bx = rem / byn, by = rem - bx * byn;
    for (int64_t t = t0; t < t1; ++t) {
        const TileCoord tc = tile_coord(cf, t);
        const int z = tc.z0 + bz, zc = z < Z ? z : Z - 1;
        uint32_t pidx[4] = {0, 0, 0, 0};
        if (p_const && active) {
            const uint32_t pz = cf.p_base + (uint32_t)__ldg(zoffP + zc);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = tc.x0 + 4 * bx + i;
                pidx[i] = pz + (uint32_t)__ldg(xoffP + (x < M ? x : M - 1));
            }
        }
        T acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = (T)0;
        for (int c = 0; c < cf.n_chunks; ++c, ++g) {
            if (g + D < G) feed.issue();
            if (!active) continue;
            mbar_wait(&full[stage], phase);
            const int k0 = c << cf.lgKC;
            const int n_s = K - k0 < KC ? K - k0 : KC;
            const T* st = stages + (size_t)stage * cf.stage_elems + lane;
            const T* sp = st + ((bz * TX + 4 * bx) << 5);
            const T* sq = st + ((cf.p_rows1 + bz * TY + 8 * by) << 5);
            const int32_t* so = soffP + k0;
            if (dbg & 1) {  // (tuning aid: no math)
            } else if (p_const) {
                switch (cf.lgKC) {
                    case 0: fma_chunk<T, 1, true>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    case 1: fma_chunk<T, 2, true>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    case 2: fma_chunk<T, 4, true>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    default: fma_chunk<T, 8, true>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                }
            } else {
                switch (cf.lgKC) {
                    case 0: fma_chunk<T, 1, false>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    case 1: fma_chunk<T, 2, false>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    case 2: fma_chunk<T, 4, false>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                    default: fma_chunk<T, 8, false>(sp, sq, cf.slab, ws_in, pidx, so, n_s, acc); break;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == cf.n_stages) {
                stage = 0;
                phase ^= 1u;
            }
        }
        if (active && z < Z && (int64_t)tc.b0 + lane < B) {
            const uint32_t zo = o_base + (uint32_t)__ldg(zoffO + z);
            T* ob = ws_out + ws_off0 + tc.b0 + lane;
            uint32_t yo[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int y = tc.y0 + 8 * by + j;
                yo[j] = (uint32_t)__ldg(yoffO + (y < N ? y : N - 1));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int x = tc.x0 + 4 * bx + i;
                if (x < M) {
                    const uint32_t xo = zo + (uint32_t)__ldg(xoffO + x);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (tc.y0 + 8 * by + j < N) ob[(size_t)(xo + yo[j]) * ldb] = acc[i][j];
                }
            }
        }
    }
}

int pow2_ceil_log(int x) {
    int l = 0;
    while ((1 << l) < x) ++l;
    return l;
}

}  // namespace

bool mm_pick(const int32_t* r, size_t item_bytes, bool allow_mma, int64_t ldb, MMChoice& out) {
    const int A = r[0], S = r[1], K2 = r[2], flags = r[3];
    const int64_t out_size = ld_i64(r + 4), sum_size = ld_i64(r + 6);
    if (K2 != 2 || flags != 0 || A < 1 || S < 1 || sum_size < 2 || sum_size > 4096 || out_size * sum_size < 2048) return false;
    const int opw = OP_FIXED + A + S;
    const int32_t* odims = r + STEP_FIXED;
    const int32_t* sdims = odims + A;
    const int32_t* ops = sdims + S;
    bool work[2];
    for (int k = 0; k < 2; ++k) {
        const int32_t* op = ops + k * opw;
        if (op[3] != 0) return false;  // evidence-dependent operand: per-lane base, rows are not warp uniform
        work[k] = (op[0] & 0xFF) == 1;
        if (ld_i64(op + 1) < 0 || ld_i64(op + 1) >= (1LL << 31)) return false;
    }
    if (!work[0] && !work[1]) return false;
    if (ld_i64(r + 8) >= (1LL << 31)) return false;
    std::vector<int64_t> ostr(A);
    {
        int64_t acc = 1;
        for (int a = A - 1; a >= 0; --a) {
            ostr[a] = acc;
            acc *= odims[a];
        }
    }
    bool found = false;
    MMChoice best;
    for (int sw = 0; sw < 2; ++sw) {
        const int32_t* P = ops + sw * opw;
        const int32_t* Q = ops + (1 - sw) * opw;
        const bool p_work = work[sw], q_work = work[1 - sw];
        if (!q_work) continue;  // a batch-invariant operand always plays P
        std::vector<int> am, an, az;
        for (int a = 0; a < A; ++a) {
            const bool ip = P[OP_FIXED + a] != 0, iq = Q[OP_FIXED + a] != 0;
            if (ip && !iq) am.push_back(a);
            else if (iq && !ip) an.push_back(a);
            else az.push_back(a);
        }
        int64_t M = 1, N = 1, Z = 1;
        for (int a : am) M *= odims[a];
        for (int a : an) N *= odims[a];
        for (int a : az) Z *= odims[a];
        const int64_t K = sum_size;
        if (M < 3 || N < 3 || M > (1 << 20) || N > (1 << 20) || Z > (1 << 22)) continue;
        // tensor cores unless the 8-row DMMA fragments would pad M by more than 15 % (M = 11, 13, 19 ...: measured per
        // launch on diabetes, those steps lose 30-40 % against the 4-row FMA blocks, while CPT steps with M = 100..221
        // gain 15-20 %; profiles/r02_mm_kernel.md)
        // (also allowing M <= 8 helps munin's M = 2..6, K = 80..100 steps by 0.6 ms but costs diabetes 3.6 ms on its
        // M = 5, N = 11, K = 221 steps: not taken)
        const bool mma = allow_mma && !p_work && item_bytes == 8 && ((M + 7) / 8 * 8) * 100 <= M * 115;
        const int row_bytes = 32 * (int)item_bytes;
        // tile search: blocks of 4 x 8 (FMA) or 8 x 4 (DMMA) outputs per warp, one block per warp
        double bc = 1e300;
        int b_lgTX = 0, b_lgTY = 0, b_TZ = 0, b_lgKC = 0, b_stages = 0;
        for (int lgTX = mma ? 3 : 2; lgTX <= 6; ++lgTX) {
            const int TX = 1 << lgTX;
            if (lgTX > (mma ? 3 : 2) && TX / 2 >= M) continue;
            for (int lgTY = mma ? 2 : 3; lgTY <= 7; ++lgTY) {
                const int TY = 1 << lgTY;
                if (lgTY > (mma ? 2 : 3) && TY / 2 >= N) continue;
                const int blocks = (TX * TY) / 32;
                if (blocks > MM_CONSUMERS) continue;
                int TZ = MM_CONSUMERS / blocks;
                if (TZ > Z) TZ = (int)Z;
                const int64_t ntx = (M + TX - 1) / TX, nty = (N + TY - 1) / TY, ntz = (Z + TZ - 1) / TZ;
                const int rows_per_k = TZ * ((p_work ? TX : 0) + TY);
                const int skew = mma ? MMA_SKEW * (int)item_bytes : 0;
                const int64_t bytes_per_k = (int64_t)rows_per_k * row_bytes + skew;
                int lgKC = pow2_ceil_log((int)std::min<int64_t>(K, 8));
                int stages = 0;
                for (; lgKC >= 0; --lgKC) {
                    stages = (int)std::min<int64_t>(MM_MAX_STAGES, (int64_t)MM_SMEM_BUDGET / (bytes_per_k << lgKC));
                    if (stages >= 3 || (stages >= 2 && lgKC == 0)) break;
                }
                if (lgKC < 0) continue;
                const int64_t n_chunks = (K + (1 << lgKC) - 1) >> lgKC;
                const double tiles = (double)ntx * nty * ntz;
                // model: shared-memory wavefronts of the consumers + L2 -> shared bytes of the stages + per-stage sync
                const double cost = tiles * ((double)K * 24.0 * blocks * TZ + (double)K * rows_per_k * 6.0 + n_chunks * 150.0 + 400.0);
                if (cost < bc) {
                    bc = cost;
                    b_lgTX = lgTX;
                    b_lgTY = lgTY;
                    b_TZ = TZ;
                    b_lgKC = lgKC;
                    b_stages = stages;
                }
            }
        }
        if (bc >= 1e299) continue;
        if (found && bc >= best.cost) continue;
        MMChoice ch;
        MMItem& it = ch.item;
        std::memset(&it, 0, sizeof(it));
        it.M = (int32_t)M;
        it.N = (int32_t)N;
        it.Z = (int32_t)Z;
        it.K = (int32_t)K;
        it.lgTX = b_lgTX;
        it.lgTY = b_lgTY;
        it.TZ = b_TZ;
        it.lgKC = b_lgKC;
        const int TX = 1 << b_lgTX, TY = 1 << b_lgTY, KC = 1 << b_lgKC;
        it.ntx = (int32_t)((M + TX - 1) / TX);
        it.nty = (int32_t)((N + TY - 1) / TY);
        it.ntz = (int32_t)((Z + b_TZ - 1) / b_TZ);
        it.n_tiles = it.ntx * it.nty * it.ntz;
        it.n_chunks = (int32_t)((K + KC - 1) / KC);
        it.p_const = p_work ? 0 : 1;
        it.use_mma = mma ? 1 : 0;
        const int64_t rows1 = (int64_t)b_TZ * ((p_work ? TX : 0) + TY);
        it.q_off = (int32_t)((p_work ? (int64_t)b_TZ * TX : 0) * 32);  // first Q row inside a summed-index slab
        it.stage_elems = (int32_t)(KC * (rows1 * 32 + (mma ? MMA_SKEW : 0)));
        it.n_stages = (int32_t)std::min<int64_t>(b_stages, std::max<int64_t>(1, (int64_t)it.n_chunks * it.n_tiles * 4));
        it.p_base = (uint32_t)ld_i64(P + 1);
        it.q_base = (uint32_t)ld_i64(Q + 1);
        it.o_base = (uint32_t)ld_i64(r + 8);
        it.n_active = b_TZ * (TX * TY / 32);
        if (const char* e = std::getenv("PGX_MM_DEBUG")) it.pad0 = std::atoi(e);  // tuning aid: 1 no math, 2 no copies
        ch.smem = 128 + (size_t)it.n_stages * it.stage_elems * item_bytes;
        ch.cost = bc;
        // offset tables (table entries): flattened M / N / Z / K index -> offset inside P, Q and the output
        auto tabulate = [&](const std::vector<int>& axes, const int32_t* dims, auto stride_of, std::vector<int32_t>& dst) {
            int64_t n = 1;
            for (int a : axes) n *= dims[a];
            const size_t at = dst.size();
            dst.resize(at + (size_t)n);
            for (int64_t i = 0; i < n; ++i) {
                int64_t rem = i, off = 0;
                for (int j = (int)axes.size() - 1; j >= 0; --j) {
                    const int a = axes[j];
                    off += (rem % dims[a]) * stride_of(a);
                    rem /= dims[a];
                }
                dst[at + (size_t)i] = (int32_t)off;
            }
        };
        std::vector<int> as(S);
        for (int a = 0; a < S; ++a) as[a] = a;
        std::vector<int32_t>& tb = ch.tabs;
        tabulate(am, odims, [&](int a) { return (int64_t)P[OP_FIXED + a]; }, tb);
        tabulate(am, odims, [&](int a) { return ostr[a]; }, tb);
        tabulate(an, odims, [&](int a) { return (int64_t)Q[OP_FIXED + a]; }, tb);
        tabulate(an, odims, [&](int a) { return ostr[a]; }, tb);
        tabulate(as, sdims, [&](int a) { return (int64_t)P[OP_FIXED + A + a]; }, tb);
        tabulate(as, sdims, [&](int a) { return (int64_t)Q[OP_FIXED + A + a]; }, tb);
        tabulate(az, odims, [&](int a) { return (int64_t)P[OP_FIXED + a]; }, tb);
        tabulate(az, odims, [&](int a) { return (int64_t)Q[OP_FIXED + a]; }, tb);
        tabulate(az, odims, [&](int a) { return ostr[a]; }, tb);
        // the summed-index offsets once more, in ELEMENTS of a work table (x ldb), for the copy loop
        tabulate(as, sdims, [&](int a) { return (int64_t)P[OP_FIXED + A + a] * (p_work ? ldb : 0); }, tb);
        tabulate(as, sdims, [&](int a) { return (int64_t)Q[OP_FIXED + A + a] * ldb; }, tb);
        best = std::move(ch);
        found = true;
    }
    if (!found) return false;
    out = std::move(best);
    return true;
}

cudaError_t mm_launch(size_t item_bytes, const MMItem* d_items, int n_items, int n_blocks, size_t smem, const int32_t* d_tabs,
                      void* ws_all, uint32_t ws_off0, int64_t B, uint32_t ldb, int b_tiles, cudaStream_t st) {
    cudaError_t e;
    if (item_bytes == 8) {
        e = cudaFuncSetAttribute(k_contract_mm<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_contract_mm<double><<<(unsigned)n_blocks, MM_THREADS, smem, st>>>(d_items, n_items, d_tabs, (const double*)ws_all,
                                                                           (double*)ws_all, ws_off0, B, ldb, b_tiles);
    } else {
        e = cudaFuncSetAttribute(k_contract_mm<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        k_contract_mm<float><<<(unsigned)n_blocks, MM_THREADS, smem, st>>>(d_items, n_items, d_tabs, (const float*)ws_all,
                                                                          (float*)ws_all, ws_off0, B, ldb, b_tiles);
    }
    return cudaGetLastError();
}

}  // namespace pgx
