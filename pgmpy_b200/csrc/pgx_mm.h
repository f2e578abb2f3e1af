// pgx_mm.h — interface between pgx.cu (schedules, C-ABI) and pgx_mm.cu (K3: matrix-product-shaped contraction steps).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>
#include <vector>

namespace pgx {

// Warps per CTA. 8 = two CTAs per SM, measured best overall (diabetes / munin, ms): 16 warps x 1 CTA 53.9 / 34.1,
// 8 x 2 53.2 / 33.0, 4 x 4 54.6 / 32.6 — smaller CTAs overlap each other's chunk boundaries, larger tiles re-read less.
#ifndef PGX_MM_WARPS
#define PGX_MM_WARPS 8
#endif
constexpr int MM_WARPS = PGX_MM_WARPS;                  // warps per CTA: each copies its share of the stage rows and owns one
constexpr int MM_CONSUMERS = MM_WARPS;                  // 4 x 8 register block of the tile (128 registers per thread)
constexpr int MM_THREADS = 32 * MM_WARPS;
constexpr int MM_CTAS_PER_SM = 16 / MM_WARPS;           // 16 warps x 128 registers fill the register file either way
constexpr int MM_MAX_STAGES = 8;
constexpr size_t MM_SMEM_BUDGET = (MM_CTAS_PER_SM == 1 ? 208 : (MM_CTAS_PER_SM == 2 ? 110 : 54)) * 1024;  // stage ring, per CTA

// One step of a k_contract_mm launch. A two-operand sum-product step
//     out[o, b] = sum_s P[ip(o, s), b] * Q[iq(o, s), b]
// is seen as Z independent matrix products out_z[M, N] = P_z[M, K] * Q_z[K, N] per evidence set:
//     M = output axes only P depends on, N = only Q, Z = both (a batch dimension), K = the summed axes,
// each flattened to one index whose table offsets the HOST tabulates once per schedule (xoff/yoff/zoff/soff below),
// so the kernel does no mixed-radix arithmetic at all.
struct MMItem {
    int32_t blk_begin;      // first CTA of this step inside the launch
    int32_t n_ctas;         // CTAs serving this step
    int32_t tiles_per_cta;  // consecutive tiles one CTA walks (a persistent producer/consumer pipeline runs across them)
    int32_t n_tiles;        // tiles per tile of 32 evidence sets = ntz * ntx * nty
    int32_t M, N, Z, K;
    int32_t lgTX, lgTY, TZ, lgKC;  // tile = TZ x 2^lgTX x 2^lgTY outputs, 2^lgKC summed indices per pipeline stage
    int32_t ntx, nty, ntz, n_chunks;
    int32_t n_stages;       // pipeline depth
    int32_t stage_elems;    // elements (of T) per stage
    int32_t q_off;          // element offset of the first Q row inside a summed-index slab of a stage (P rows first)
    int32_t p_const;        // 1: P is a batch-invariant table (read as scalars from the table copy, never staged)
    uint32_t p_base, q_base, o_base;  // table offsets in entries (const: index into the table copy; work: workspace entry)
    int32_t tab;            // word offset of this step's offset tables inside the launch's table pool:
                            // xoffP[M] xoffO[M] yoffQ[N] yoffO[N] soffP[K] soffQ[K] zoffP[Z] zoffQ[Z] zoffO[Z] (entries),
                            // then soffP[K] * ldb, soffQ[K] * ldb (elements, for the copy loop)
    int32_t n_active;       // consumer warps with a register block = TZ * (TX/4) * (TY/8)
    int32_t use_mma;        // 1: fp64 tensor-core (DMMA m8n8k4) consumer for a batch-invariant P
    int32_t pad0, pad1;
};

struct MMChoice {
    MMItem item;
    std::vector<int32_t> tabs;
    size_t smem = 0;
    double cost = 0;  // model cycles per tile of 32 evidence sets (for the eligibility decision)
};

// Host: is this step record (layout in pgmpy_b200/plan.py) matrix-product shaped, and with which tiling?
bool mm_pick(const int32_t* rec, size_t item_bytes, bool allow_mma, int64_t ldb, MMChoice& out);

cudaError_t mm_launch(size_t item_bytes, const MMItem* d_items, int n_items, int n_blocks, size_t smem, const int32_t* d_tabs,
                      void* ws_all, uint32_t ws_off0, int64_t B, uint32_t ldb, int b_tiles, cudaStream_t st);

// fp32 mode, batch-invariant P: the same step on the tcgen05 tensor cores as a TF32x3 GEMM with TMEM accumulators
// (pgx_tc32.cu). Uses the MMItem / offset tables of mm_pick; tiles_per_cta = n per CTA, one CTA per (z, 128 evidence
// sets, run of n).
bool tc32_eligible(const MMItem& it);
size_t tc32_smem_bytes(const MMItem& it);
cudaError_t tc32_launch(const MMItem* d_items, int n_items, int n_blocks, size_t smem, const int32_t* d_tabs, void* ws_all,
                        uint32_t ws_off0, int64_t B, uint32_t ldb, int b_chunks, cudaStream_t st);

}  // namespace pgx
