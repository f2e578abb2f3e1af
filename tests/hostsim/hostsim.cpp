// hostsim.cpp — TEST INFRASTRUCTURE. Compiles the device element function of pgmpy_b200/csrc/pgx_step.cuh with g++
// and walks a packed plan on the CPU exactly as the kernels do (same step records, same [entry][ldb] layout),
// so index arithmetic can be checked against oracle/plan_exec.py without a GPU. Never linked into libpgx.so.
#include <cstdint>
#include <vector>

#include "../../pgmpy_b200/csrc/pgx_step.cuh"

using namespace pgx;

template <typename T>
static void run(const int32_t* pool, const T* cst, const int32_t* ev, T* ws, T* out, int64_t B, int64_t ldb) {
    const int n_ev = pool[2], n_steps = pool[3], n_segs = pool[4];
    const int64_t out_elems = pool[5];
    const int32_t* index = pool + pool[10];
    const int32_t* ev_card = pool + pool[14];
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* rec = pool + index[s];
        const uint32_t out_size = (uint32_t)rec[4];
        const int64_t out_off = ld_i64(rec + 8);
        for (uint32_t o = 0; o < out_size; ++o)
            for (int64_t b = 0; b < B; ++b)
                ws[(out_off + o) * ldb + b] = contract_elem_upto<T, MAX_OPS>(rec, cst, ws, ev + b * n_ev, ev_card, ldb, b, o);
    }
    const int32_t* segs = pool + pool[11];
    for (int g = 0; g < n_segs; ++g) {
        const int32_t* seg = segs + g * SEG_WORDS;
        const int64_t off = ld_i64(seg);
        for (int64_t b = 0; b < B; ++b) {
            T sum = 0;
            for (int i = 0; i < seg[2]; ++i) sum += ws[(off + i) * ldb + b];
            for (int i = 0; i < seg[2]; ++i) {
                const T v = ws[(off + i) * ldb + b];
                out[b * out_elems + seg[3] + i] = (seg[4] & SEG_NORMALIZE) ? v / sum : v;
            }
        }
    }
}

extern "C" void hostsim_run_f64(const int32_t* pool, const double* cst, const int32_t* ev, double* ws, double* out,
                                int64_t B, int64_t ldb) {
    run<double>(pool, cst, ev, ws, out, B, ldb);
}
extern "C" void hostsim_run_f32(const int32_t* pool, const float* cst, const int32_t* ev, float* ws, float* out, int64_t B,
                                int64_t ldb) {
    run<float>(pool, cst, ev, ws, out, B, ldb);
}
