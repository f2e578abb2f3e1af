/*
 * pgx.h — C-ABI of libpgx.so, the B200 (sm_100a) exact-inference engine behind pgmpy's
 * VariableElimination.query / BeliefPropagation.calibrate,query / DiscreteFactor algebra.
 *
 * The reference (tristantreb/pgmpy v1.0.0) is pure Python and has NO plugin/operator/FFI interface
 * (SURVEY.md §8b): its only seam is the per-op function table pgmpy/utils/compat_fns.py:21-180 selected by
 * pgmpy/global_vars.py:82-128. The entry points below are therefore what a binding for this path
 * would call instead of that table: one compiled plan per (model, query signature), executed over a
 * batch of independent evidence sets.
 *
 * What each entry point replaces in the reference (paths relative to /root/reference/pgmpy/):
 *   pgx_plan_create      inference/base.py:88-152 (_initialize_structures: per-variable factor index) and
 *                        the per-query control flow of inference/ExactInference.py:141-244 (VE loop) /
 *                        :854-895 (junction-tree calibration schedule), frozen into a static step list.
 *   pgx_run_batch        the arithmetic of ExactInference.py:200-229 and :770-805, i.e. per step
 *                          factors/discrete/DiscreteFactor.py:599-614  reduce   (evidence index gather)
 *                          factors/discrete/DiscreteFactor.py:769-777  product  (einsum outer join)
 *                          factors/discrete/DiscreteFactor.py:400-408  marginalize (einsum sum-out)
 *                          factors/discrete/DiscreteFactor.py:838-863  divide   (0/0 -> 0, x/0 -> inf)
 *                          factors/discrete/DiscreteFactor.py:530      normalize (NaN when the sum is 0)
 *                        for B evidence sets at once.
 *   pgx_evidence_reduce  DiscreteFactor.reduce (:535-617) / the greedy-path indexer ExactInference.py:355-365.
 *   pgx_normalize        DiscreteFactor.normalize (:485-533).
 *   pgx_argmax_rows      compat_fns.argmax over the joint in map_query (inference/ExactInference.py:611).
 *
 * Conventions: plain pointers and sizes only; 0 = success, negative = error (pgx_last_error() gives the
 * text, thread local); no exceptions cross the boundary. The CALLER owns every device buffer
 * (table blob, evidence, output, workspace); a plan owns only device copies of its descriptors.
 * A plan may be enqueued on several streams with different workspaces (the end-to-end ring of pgmpy_b200.engine does
 * that); calls on the SAME plan from several host threads must be serialised by the caller (launch schedules and CUDA
 * graphs are cached inside the plan).
 */
#ifndef PGX_H
#define PGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGX_ABI_VERSION 1

/* dtype of the table blob, the workspace and the output */
#define PGX_F64 0
#define PGX_F32 1

/* error codes */
#define PGX_OK 0
#define PGX_ERR_INVALID (-1)   /* malformed descriptor / argument */
#define PGX_ERR_BOUNDS (-2)    /* a step would read or write outside its tables */
#define PGX_ERR_WORKSPACE (-3) /* workspace too small */
#define PGX_ERR_CUDA (-4)      /* CUDA runtime error (text in pgx_last_error) */
#define PGX_ERR_UNSUPPORTED (-5)

/* execution modes (pgx_plan_set_option(PGX_OPT_MODE, ...)) */
#define PGX_MODE_AUTO 0     /* pick by plan size and batch */
#define PGX_MODE_STEPWISE 1 /* one launch per step, grid over (output entries x evidence sets) */
#define PGX_MODE_FUSED 2    /* whole plan in one launch, a warp-row of 32 evidence sets per CTA */

#define PGX_OPT_MODE 1
#define PGX_OPT_FUSED_WARPS 2 /* warps cooperating on one row of 32 evidence sets in fused mode (1..32) */
#define PGX_OPT_USE_GRAPH 3   /* stepwise mode: replay the launch sequence as a CUDA graph (0/1) */
#define PGX_OPT_FUSED_KERNEL 4 /* 0 auto | 1 generic addressing | 2 offset tables + shared-memory work tables |
                                  3 offset tables + global work tables | 4 the plan-specialised kernel
                                  (pgx_plan_specialize; auto uses it whenever it has been built) */

#define PGX_OPT_STEP_KERNEL 5  /* stepwise mode: 0 auto (tile-cooperative kernel, 32-bit addressing) | 1 generic only |
                                  2 tile-cooperative kernel with 64-bit addressing */

/* options 6, 7, 8 (register-tiled / 2-D register-tiled / two-sets-per-lane variants of the step kernel) were measured
 * slower than the default on every model (profiles/r01_final_summary.md) and are retired; the numbers stay reserved. */

#define PGX_OPT_STAGE 9        /* stepwise mode, matrix-product-shaped two-operand steps: 1 (default) the pipelined
                                  shared-memory-staged tile kernel k_contract_mm (pgx_mm.cu) | 0 the streaming kernel only */
#define PGX_OPT_MMA 10         /* k_contract_mm: fp64 tensor cores (DMMA m8n8k4) for steps whose first operand is a
                                  batch-invariant table, [M x K] . [K x (N . B)] (default 1; 0 = FMA consumers) */

#define PGX_OPT_TC32 11        /* fp32 mode: steps whose first operand is a batch-invariant table with >= 32 rows run as a
                                  TF32x3 GEMM on the tcgen05 tensor cores with TMEM accumulators (pgx_tc32.cu; default 1) */

#define PGX_INFO_N_STEPS 1
#define PGX_INFO_OUT_ELEMS 2
#define PGX_INFO_WS_ENTRIES 3
#define PGX_INFO_LAST_LAUNCHES 4 /* kernels launched by the most recent pgx_run_batch */
#define PGX_INFO_LAST_MODE 5
#define PGX_INFO_N_EV 6
#define PGX_INFO_LAST_VARIANT 7 /* which fused kernel ran (PGX_OPT_FUSED_KERNEL numbering), 0 if stepwise */
#define PGX_INFO_LAST_GRAPH 9    /* 1 if the most recent stepwise run was a CUDA-graph replay */
#define PGX_INFO_LAST_STAGED_STEPS 10 /* steps the most recent stepwise run sent to the TMA-staged GEMM-tile kernel */
#define PGX_INFO_LAST_TC_STEPS 12 /* steps the most recent stepwise run sent to the tcgen05 kernel (fp32 mode) */
#define PGX_INFO_IN_ELEMS 11    /* elements of one row of `soft` (0: the plan has no input tables) */
#define PGX_INFO_N_LEVELS 8     /* dependency levels of the plan (0 when no offset tables were built) */
#define PGX_INFO_SPECIALIZED 13 /* 1 after a successful pgx_plan_specialize */
#define PGX_INFO_SPEC_REGS 14   /* registers per thread / dynamic shared memory / NVRTC+load time of the specialised kernel */
#define PGX_INFO_SPEC_SMEM 15
#define PGX_INFO_SPEC_COMPILE_MS 16
#define PGX_INFO_SPEC_LOADS 17  /* loads and fp instructions the generator emitted per row of 32 evidence sets */
#define PGX_INFO_SPEC_FLOPS 18

typedef struct pgx_plan pgx_plan; /* opaque */

typedef struct pgx_plan_desc {
    int32_t abi_version;      /* PGX_ABI_VERSION */
    int32_t dtype;            /* PGX_F64 | PGX_F32 */
    const int32_t* pool;      /* HOST: plan word pool, layout in pgmpy_b200/plan.py (header, evidence cards,
                                 step index, step records, output segments) */
    int64_t pool_words;
    const void* table_blob;   /* DEVICE: packed batch-invariant tables (CPTs / clique potentials), dtype above */
    int64_t table_entries;    /* entries in table_blob */
} pgx_plan_desc;

/* Validates the pool (every operand range is bounds-checked against table_entries / the workspace size)
 * and uploads the descriptors to the current CUDA device. */
int pgx_plan_create(const pgx_plan_desc* desc, pgx_plan** out);
void pgx_plan_destroy(pgx_plan* plan);

/* Bytes of workspace pgx_run_batch needs for a batch of B evidence sets with the plan's CURRENT options: 0 once the
 * plan has been specialised (pgx_plan_specialize) and no other kernel is forced — `workspace` may then be NULL. */
size_t pgx_workspace_bytes(const pgx_plan* plan, int64_t B);

/* Runs the plan for B evidence sets on `stream` (a cudaStream_t; NULL = default stream).
 *   ev_states  DEVICE int32 [B, n_ev]  state index of each evidence variable (slot order of the plan);
 *                                      may be NULL when the plan has no evidence slots
 *   out        DEVICE dtype [B, out_elems] posterior rows (segments normalised as the plan says)
 * Asynchronous: returns after enqueueing. */
int pgx_run_batch(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                  int64_t B, void* stream);

/* n independent pgx_run_batch calls enqueued on `stream` by ONE call across the boundary: a mixed-evidence batch
 * (pgmpy/models/DiscreteBayesianNetwork.py:973-989 feeds the reference one row at a time; here the rows are bucketed
 * by observed set, one plan per bucket) is a handful of short launches, and the host path of a call per bucket — not
 * the kernels — was what bounded it. Arrays of n entries; stops at the first error ("job i: ..."). */
int pgx_run_batch_multi(int32_t n, pgx_plan* const* plans, const int32_t* const* ev_states, void* const* outs,
                        void* const* workspaces, const size_t* workspace_bytes, const int64_t* B, void* stream);

/* The same for a plan with batch-dependent INPUT tables (soft / virtual evidence, pgmpy/inference/base.py:256-299:
 * one likelihood vector per soft-evidence variable, which the reference adds as an observed binary child per query):
 *   soft       DEVICE dtype [B, in_elems]   row b = the input tables of evidence set b (PGX_INFO_IN_ELEMS, layout in
 *                                           the plan's inputs block, pgmpy_b200/plan.py)
 * pgx_run_batch on such a plan is an error; `soft` must be NULL for a plan without input tables. */
int pgx_run_batch_soft(pgx_plan* plan, const int32_t* ev_states, const void* soft, void* out, void* workspace,
                       size_t workspace_bytes, int64_t B, void* stream);

/* Max-product with back-pointers (most probable explanation over all unobserved variables; the reference maximises the
 * full joint table, pgmpy/inference/ExactInference.py:609-612, :1222-1317). pgx_plan_set_trace attaches the traceback
 * descriptor of a max-product junction-tree plan (HOST int32 words, layout in pgmpy_b200/planner.py
 * compile_jt_mpe_plan: per clique, root first, the table of its upward belief and which of its axes are assigned
 * there); pgx_run_batch_mpe runs the plan's steps and then the traceback kernel:
 *   assign     DEVICE int32 [B, n_columns]  state index of every unobserved variable (column order of the descriptor)
 * `soft` as in pgx_run_batch_soft (NULL when the plan has no input tables). */
int pgx_plan_set_trace(pgx_plan* plan, const int32_t* trace, int64_t n_words);
int pgx_run_batch_mpe(pgx_plan* plan, const int32_t* ev_states, const void* soft, int32_t* assign, void* workspace,
                      size_t workspace_bytes, int64_t B, void* stream);

/* Tracing aid: runs the plan in stepwise mode with a CUDA event after every step and returns the per-step device
 * time in milliseconds (step_ms[n_steps], n_steps >= the plan's step count). Synchronises the stream. */
int pgx_profile_steps(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                      int64_t B, void* stream, float* step_ms, int32_t n_steps);

/* Tracing aid: the same pass with the production launch schedule (steps of one dependency level share a launch).
 * launch_ms[i] = device time of launch i (cap_launches >= number of launches), step_launch[s] = launch that served
 * plan step s (n_steps >= the plan's step count), *n_launches = launches recorded. Synchronises the stream. */
int pgx_profile_launches(pgx_plan* plan, const int32_t* ev_states, void* out, void* workspace, size_t workspace_bytes,
                         int64_t B, void* stream, float* launch_ms, int32_t cap_launches, int32_t* step_launch,
                         int32_t n_steps, int32_t* n_launches);

/* Host only (no GPU needed): how k_contract_mm (pgx_mm.cu) would run one step record: the step seen as Z matrix
 * products out_z[M, N] = P_z[M, K] Q_z[K, N] per evidence set. fields[24] = eligible, M, N, Z, K, lgTX, lgTY, TZ, lgKC,
 * ntx, nty, ntz, n_chunks, n_stages, stage_elems, q_off, p_const, p_base, q_base, o_base, n_active, use_mma,
 * smem_bytes, n_tiles. tabs (may be NULL; *n_tabs = words needed) receives the offset tables the kernel indexes:
 * xoffP[M] xoffO[M] yoffQ[N] yoffO[N] soffP[K] soffQ[K] zoffP[Z] zoffQ[Z] zoffO[Z] (table entries), then soffP[K] and
 * soffQ[K] again in work-table elements (x ldb; ldb = 1 in this host-only call, and 0 for a batch-invariant P). */
int pgx_mm_pick(const int32_t* step_record, int32_t item_bytes, int32_t allow_mma, int32_t* fields, int32_t* tabs,
                int64_t tabs_cap, int64_t* n_tabs);

/* Plan-specialised whole-plan kernel (pgx_spec.cu). The per-query control flow the reference interprets in Python on
 * every call (ExactInference.py:141-244, :770-895) is, for one evidence signature, a fixed list of multiply-adds:
 * pgx_plan_specialize writes it out as straight-line CUDA C for sm_100a (one warp = 32 evidence sets, work tables in
 * the warp's shared memory, batch-invariant CPT entries as immediates), compiles it with NVRTC (dlopen'ed libnvrtc;
 * PGX_ERR_UNSUPPORTED when absent or when the plan has divide / max steps, input tables or more than 120 000 product
 * terms) and loads it; later pgx_run_batch calls use it (PGX_INFO_LAST_VARIANT = 4). The table blob must not change
 * afterwards. Costs seconds: worth it from ~10^8 evidence sets per plan. With PGX_SPEC_CACHE_DIR set, compiled kernels
 * are kept there under the hash of their source and a later process pays a file read instead of the compile.
 * pgx_spec_source: host only, no GPU: the generated source (compile = 0), or compile it too (1; 2 = return the cubin
 * instead of the source). desc->table_blob is a HOST pointer here. Returns the byte count (buf receives at most cap
 * bytes), or < 0 with the reason in buf. stats8 = product terms, terms kept (non-zero coefficient), loads, fp
 * instructions, work entries, shared-memory bytes, compile ms, cubin bytes. */
int pgx_plan_specialize(pgx_plan* plan);
int64_t pgx_spec_source(const pgx_plan_desc* desc, int32_t compile, char* buf, int64_t cap, int64_t* stats8);

int pgx_plan_set_option(pgx_plan* plan, int32_t option, int64_t value);
int pgx_plan_get_info(const pgx_plan* plan, int32_t what, int64_t* value);

/* Stand-alone batched evidence reduce (index gather):
 *   dst[e, b] = table[ base(e) + sum_j ev_states[b, slot_j] * ev_stride_j ],  e over the free axes (row-major),
 *   base(e) = sum_a digit_a(e) * free_stride_a.  dst layout [free entries][ldb], b fastest. */
int pgx_evidence_reduce(int32_t dtype, const void* table, int64_t table_entries, int32_t n_free,
                        const int32_t* free_dims, const int32_t* free_strides, int32_t n_ev, const int32_t* ev_slots,
                        const int32_t* ev_strides, const int32_t* ev_cards, const int32_t* ev_states, int32_t ev_row_len,
                        void* dst, int64_t B, int64_t ldb, void* stream);

/* Stand-alone batched normalise: out[b, i] = src[i, b] / sum_i src[i, b]  (src layout [n][ldb]). */
int pgx_normalize(int32_t dtype, const void* src, int64_t n, int64_t ldb, void* out, int64_t out_row_len, int64_t B,
                  void* stream);

/* Row-wise argmax of a [B, n] row-major table (first maximum, like numpy.argmax): the decode step of map_query
 * (pgmpy/inference/ExactInference.py:611-612: argmax over the joint, then DiscreteFactor.assignment). */
int pgx_argmax_rows(int32_t dtype, const void* src, int64_t n, int64_t B, int32_t* out, void* stream);

/* Leading dimension (in evidence sets) of workspace tables for a batch of B. */
int64_t pgx_batch_ld(int64_t B);

const char* pgx_last_error(void);
int32_t pgx_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* PGX_H */
