// pgx_tc32.cu — K3 on the 5th-generation tensor cores for fp32 mode: CPT-times-message steps as a TF32x3 GEMM with TMEM
// accumulators (tcgen05.mma kind::tf32, sm_100a).
//
//     out[m, n, z, b] = sum_k C[m, k, z] * W[k, n, z, b]          C batch invariant (a CPT), W a message, b fastest
//
// With the evidence batch folded into the GEMM's row dimension this is, per (n, z), D[128 b x M] = A[128 b x K] . B[K x M]:
//   A = W^T: 128 consecutive evidence sets (the MMA's M = 128 rows = 128 TMEM lanes) x the summed indices,
//   B = C  : the CPT, staged once per z, MMA N = M rounded up to 16 (<= 256 TMEM columns), K = 8 per instruction.
// fp32 accuracy comes from the 3xTF32 split: x = hi + lo with hi = x truncated to 10 mantissa bits (exactly representable
// in TF32) and lo = x - hi (exact in fp32), D += A_hi B_hi + A_lo B_hi + A_hi B_lo, fp32 accumulation in TMEM; the dropped
// lo.lo term is 2^-20 relative. One CTA = 128 threads: thread r owns evidence set b0 + r: it writes row r of A (hi and
// lo, K-major canonical layout: 8 x 16-byte core matrices, no swizzle) and, after thread 0 has issued the MMAs and their
// tcgen05.commit has arrived on an mbarrier, reads TMEM lane r (tcgen05.ld 32x32b) and stores out[m, n, z, b0 + r] for
// every m — 128-byte coalesced segments per warp. fp64 has no tcgen05 kind (DMMA, pgx_mm.cu); this path exists for
// fp32 mode only and is selected by shape (tc32_eligible: batch-invariant P with M >= 64, K <= 64; pgx.cu: B >= 128).
#include <cstdint>

#include "pgx_mm.h"

namespace pgx {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no swizzle: element (row, k) of an operand with Kp columns sits at
//   (row / 8) * SBO + (k / 4) * 128 + (row % 8) * 16 + (k % 4) * 4   bytes,  SBO = Kp / 4 * 128
// (core matrix = 8 rows x 16 bytes; LBO = 128 B between the two core matrices an instruction's K = 8 spans).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address
    d |= (uint64_t)((128u >> 4) & 0x3FFF) << 16;     // leading-dimension byte offset (K direction)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride-dimension byte offset (next 8 rows)
    d |= (uint64_t)1 << 46;                          // descriptor version (sm_100)
    return d;                                         // layout type 0 = no swizzle, base offset 0
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p; \n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32"
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// One CTA: a run of `tiles_per_cta` consecutive n for one (z, chunk of 128 evidence sets).
__global__ void __launch_bounds__(128) k_contract_tc32(const MMItem* __restrict__ items, int n_items,
                                                       const int32_t* __restrict__ tabs, const float* __restrict__ ws_in,
                                                       float* __restrict__ ws_out, uint32_t ws_off0, int64_t B, uint32_t ldb,
                                                       int b_chunks) {
    extern __shared__ __align__(1024) unsigned char s_raw[];
    int lo = 0, hi = n_items - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(&items[mid].blk_begin) <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const MMItem* __restrict__ it = items + lo;
    const int M = it->M, N = it->N, Z = it->Z, K = it->K;
    const int Kp = (K + 7) & ~7, Np = (M + 15) & ~15;  // MMA K per instruction = 8, MMA N a multiple of 16 (MMA M = 128)
    const int npc = it->tiles_per_cta;                // n per CTA
    const int n_runs = (N + npc - 1) / npc;
    int u = (int)blockIdx.x - it->blk_begin;          // unit -> (z, chunk of 128 evidence sets, run of n)
    const int run = u % n_runs;
    u /= n_runs;
    const int bc = u % b_chunks;
    const int z = u / b_chunks;
    const int32_t* xoffP = tabs + it->tab;
    const int32_t* xoffO = xoffP + M;
    const int32_t* yoffQ = xoffO + M;
    const int32_t* yoffO = yoffQ + N;
    const int32_t* soffP = yoffO + N;
    const int32_t* soffQ = soffP + K;
    const int32_t* zoffP = soffQ + K;
    const int32_t* zoffQ = zoffP + Z;
    const int32_t* zoffO = zoffQ + Z;

    uint64_t* mbar = reinterpret_cast<uint64_t*>(s_raw);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_raw + 8);
    unsigned char* sB_hi = s_raw + 1024;
    const uint32_t b_bytes = (uint32_t)Np * Kp * 4u, a_bytes = 128u * Kp * 4u;
    unsigned char* sB_lo = sB_hi + b_bytes;
    unsigned char* sA_hi = sB_lo + b_bytes;
    unsigned char* sA_lo = sA_hi + a_bytes;
    const uint32_t sbo = (uint32_t)(Kp / 4) * 128u;

    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t n_cols = 32;
    while ((int)n_cols < Np) n_cols <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(n_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(1u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the CPT slice of this z, split into hi / lo, K-major (zero padded to Np x Kp)
    const uint32_t pz = it->p_base + (uint32_t)__ldg(zoffP + z);
    for (int i = tid; i < Np * Kp; i += 128) {
        const int m = i / Kp, k = i - m * Kp;
        float v = 0.f;
        if (m < M && k < K) v = __ldg(ws_in + pz + (uint32_t)(__ldg(xoffP + m) + __ldg(soffP + k)));
        const float h = tf32_hi(v);
        const uint32_t off = (uint32_t)(m >> 3) * sbo + (uint32_t)(k >> 2) * 128u + (uint32_t)(m & 7) * 16u + (uint32_t)(k & 3) * 4u;
        *reinterpret_cast<float*>(sB_hi + off) = h;
        *reinterpret_cast<float*>(sB_lo + off) = v - h;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = *s_tmem;

    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = Np, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Np >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t b0 = (uint32_t)bc * 128u;
    const uint32_t b = b0 + (uint32_t)tid;
    const bool b_in = b < ldb;  // rows hold ldb elements: never read past them
    const uint32_t qz = it->q_base + (uint32_t)__ldg(zoffQ + z);
    const uint32_t oz = it->o_base + (uint32_t)__ldg(zoffO + z);
    const uint32_t a_row = (uint32_t)(tid >> 3) * sbo + (uint32_t)(tid & 7) * 16u;
    uint32_t phase = 0;
    const int n_begin = run * npc, n_end = n_begin + npc < N ? n_begin + npc : N;
    for (int n = n_begin; n < n_end; ++n) {
        // A: row tid = W[., n, z, b0 + tid], hi and lo
        const uint32_t qn = qz + (uint32_t)__ldg(yoffQ + n);
        for (int kc = 0; kc < Kp; kc += 4) {
            float w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kc + j;
                w[j] = (k < K && b_in) ? ws_in[(size_t)ws_off0 + (size_t)(qn + (uint32_t)__ldg(soffQ + k)) * ldb + b] : 0.f;
            }
            float4 vh, vl;
            vh.x = tf32_hi(w[0]), vh.y = tf32_hi(w[1]), vh.z = tf32_hi(w[2]), vh.w = tf32_hi(w[3]);
            vl.x = w[0] - vh.x, vl.y = w[1] - vh.y, vl.z = w[2] - vh.z, vl.w = w[3] - vh.w;
            const uint32_t off = a_row + (uint32_t)(kc >> 2) * 128u;
            *reinterpret_cast<float4*>(sA_hi + off) = vh;
            *reinterpret_cast<float4*>(sA_lo + off) = vl;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        __syncthreads();
        if (tid == 0) {
            for (int kb = 0; kb < Kp / 8; ++kb) {
                const uint32_t ko = (uint32_t)kb * 256u;  // two core matrices along K per instruction
                const uint64_t dah = umma_desc(smem_u32(sA_hi) + ko, sbo), dal = umma_desc(smem_u32(sA_lo) + ko, sbo);
                const uint64_t dbh = umma_desc(smem_u32(sB_hi) + ko, sbo), dbl = umma_desc(smem_u32(sB_lo) + ko, sbo);
                umma_tf32(tmem_d, dah, dbh, idesc, kb > 0 ? 1u : 0u);
                umma_tf32(tmem_d, dal, dbh, idesc, 1u);
                umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(mbar)) : "memory");
        }
        {  // wait for the MMAs (bounded: a descriptor mistake must not hang the GPU)
            uint32_t ok = 0;
            for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(ok)
                    : "r"(smem_u32(mbar)), "r"(phase)
                    : "memory");
            if (!ok) __trap();
        }
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // epilogue: TMEM lane tid = evidence set b, columns = m
        const uint32_t on = oz + (uint32_t)__ldg(yoffO + n);
        const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < Np; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + (uint32_t)c0, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (b < (uint32_t)B) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int m = c0 + j;
                    if (m < M) ws_out[(size_t)ws_off0 + (size_t)(on + (uint32_t)__ldg(xoffO + m)) * ldb + b] = __uint_as_float(v[j]);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();  // every lane of D has been read and A may be overwritten
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(n_cols));
}

}  // namespace

bool tc32_eligible(const MMItem& it) {
    const int Kp = (it.K + 7) & ~7, Np = (it.M + 15) & ~15;
    // (measured, diabetes / munin fp32 whole pass: 36.1 / 22.8 ms on FFMA, 37.2 / 22.6 ms with every M >= 32 step here —
    // these products are bound by HBM and instruction issue, not by the math pipe, so the tensor path is time neutral)
    return it.p_const && it.M >= 64 && Np <= 256 && it.K <= 64 && (size_t)Np * Kp <= 8192;
}

size_t tc32_smem_bytes(const MMItem& it) {
    const int Kp = (it.K + 7) & ~7, Np = (it.M + 15) & ~15;
    return 1024 + 2 * (size_t)Np * Kp * 4 + 2 * (size_t)128 * Kp * 4;
}

cudaError_t tc32_launch(const MMItem* d_items, int n_items, int n_blocks, size_t smem, const int32_t* d_tabs, void* ws_all,
                        uint32_t ws_off0, int64_t B, uint32_t ldb, int b_chunks, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(k_contract_tc32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    k_contract_tc32<<<(unsigned)n_blocks, 128, smem, st>>>(d_items, n_items, d_tabs, (const float*)ws_all, (float*)ws_all, ws_off0, B,
                                                           ldb, b_chunks);
    return cudaGetLastError();
}

}  // namespace pgx
