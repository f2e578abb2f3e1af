// pgx_spec.h — plan-specialised whole-plan kernel ("K5 specialised"): interface between pgx.cu and pgx_spec.cu.
//
// A junction-tree plan is a fixed list of a few thousand multiply-adds whose operand addresses are all known when the
// plan is created. pgx_spec_build writes that list out as straight-line CUDA C (one warp = 32 evidence sets, lane =
// evidence set, work entries in registers, no barrier anywhere), compiles it for sm_100a with NVRTC and loads it
// through the driver. Against the table-driven kernel (pgx_fused.cuh) this removes every offset-table load and index
// addition, keeps every message in a register instead of shared memory, folds batch-invariant CPT entries into
// immediates (terms with a zero entry vanish) and needs no level barrier; the posterior rows leave by TMA bulk store.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>

#ifndef PGX_SPEC_DEFAULT_ACC
#define PGX_SPEC_DEFAULT_ACC 1  // partial sums per output entry
#endif
#ifndef PGX_SPEC_DEFAULT_ROWS
#define PGX_SPEC_DEFAULT_ROWS 1
#endif
#ifndef PGX_SPEC_DEFAULT_WARPS
#define PGX_SPEC_DEFAULT_WARPS 1  // warps sharing one row of 32 evidence sets (measured: profiles/r02_spec_kernel.md)
#endif

namespace pgx {

struct SpecKernel;  // opaque: loaded module + launch geometry

struct SpecStats {
    int64_t terms = 0;        // product terms of the plan (out x sum over the steps)
    int64_t terms_kept = 0;   // after dropping terms with a zero batch-invariant factor
    int64_t loads = 0;        // shared/global loads emitted (distinct elements per group)
    int64_t flops = 0;        // multiply / fma / add instructions emitted
    int64_t ws_entries = 0;   // work-table entries per evidence set after lifetime packing
    int64_t smem_bytes = 0;   // dynamic shared memory per CTA
    int warps = 1;            // warps sharing one row of 32 evidence sets
    int rows = 1;             // independent rows per CTA
    int persistent = 0;       // > 0: persistent kernel, this many CTAs per SM walk the rows
    double compile_s = 0.0;
    int regs = 0;
};

// Why a plan cannot be specialised is returned in `why` (empty when it can).
// host_blob: HOST copy of the batch-invariant tables (dtype of the plan).
bool pgx_spec_generate(const int32_t* pool, int64_t pool_words, const void* host_blob, int64_t table_entries, int dtype,
                       std::string& source, SpecStats& stats, std::string& why, int warps);

// generate + NVRTC (sm_100a) + load on the current device. Returns nullptr and sets `why` on failure.
SpecKernel* pgx_spec_build(const int32_t* pool, int64_t pool_words, const void* host_blob, int64_t table_entries,
                           int dtype, std::string& why, int warps);
// generate + NVRTC only (no GPU needed): the cubin, for tests and SASS inspection
bool pgx_spec_compile(const std::string& source, std::string& cubin, std::string& log);

const SpecStats& pgx_spec_stats(const SpecKernel* k);
// Enqueue on `stream` (a cudaStream_t). cst = DEVICE table blob, ev = DEVICE int32 [B, n_ev], out = DEVICE [B, out_elems].
int pgx_spec_launch(SpecKernel* k, const void* cst, const int32_t* ev, void* out, int64_t B, void* stream, std::string& err);
void pgx_spec_destroy(SpecKernel* k);

}  // namespace pgx
