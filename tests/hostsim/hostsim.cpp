// hostsim.cpp — TEST INFRASTRUCTURE. Compiles the device element function of pgmpy_b200/csrc/pgx_step.cuh with g++
// and walks a packed plan on the CPU exactly as the kernels do (same step records, same [entry][ldb] layout),
// so index arithmetic can be checked against oracle/plan_exec.py without a GPU. Never linked into libpgx.so.
#include <cstdint>
#include <vector>

#include "../../pgmpy_b200/csrc/pgx_step.cuh"

using namespace pgx;

template <typename T>
static void run(const int32_t* pool, const T* cst, const int32_t* ev, T* ws, T* out, int64_t B, int64_t ldb, bool use_run) {
    const int n_ev = pool[2], n_steps = pool[3], n_segs = pool[4];
    const int64_t out_elems = pool[5];
    const int32_t* index = pool + pool[10];
    const int32_t* ev_card = pool + pool[14];
    for (int s = 0; s < n_steps; ++s) {
        const int32_t* rec = pool + index[s];
        const uint32_t out_size = (uint32_t)rec[4];
        const int64_t out_off = ld_i64(rec + 8);
        if (use_run && rec[2] <= 8) {
            // the stepwise kernel's path: runs of consecutive entries with incremental offsets (run length 5)
            for (int64_t b = 0; b < B; ++b)
                for (uint32_t o0 = 0; o0 < out_size; o0 += 5)
                    contract_run<T, 8>(rec, cst, ws, ev + b * n_ev, ev_card, ldb, b, o0, o0 + 5 < out_size ? o0 + 5 : out_size,
                                       ws + out_off * ldb + b);
        } else {
            for (uint32_t o = 0; o < out_size; ++o)
                for (int64_t b = 0; b < B; ++b)
                    ws[(out_off + o) * ldb + b] = contract_elem_upto<T, MAX_OPS>(rec, cst, ws, ev + b * n_ev, ev_card, ldb, b, o);
        }
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
                                int64_t B, int64_t ldb, int use_run) {
    run<double>(pool, cst, ev, ws, out, B, ldb, use_run != 0);
}
extern "C" void hostsim_run_f32(const int32_t* pool, const float* cst, const int32_t* ev, float* ws, float* out, int64_t B,
                                int64_t ldb, int use_run) {
    run<float>(pool, cst, ev, ws, out, B, ldb, use_run != 0);
}

// ---- microprogram (pgx_fused.cuh::build_micro) walked on the CPU: checks the offset tables, the level
// ---- partition and the output column map that the table-driven fused kernel consumes.
#include "../../pgmpy_b200/csrc/pgx_fused.cuh"

extern "C" int hostsim_micro_f64(const int32_t* pool, const double* cst, const int32_t* ev, double* ws, double* out,
                                 int64_t B, int64_t ldb, int32_t* n_levels_out) {
    std::vector<int32_t> mp;
    MicroInfo info;
    if (!build_micro(pool, mp, info)) return -1;
    const int n_ev = pool[2];
    const int32_t* ev_card = pool + pool[14];
    const int32_t* levels = mp.data() + mp[2];
    const int32_t* items = mp.data() + mp[3];
    for (int lv = 0; lv < mp[0]; ++lv) {
        for (int i = levels[lv]; i < levels[lv + 1]; ++i) {
            const int32_t* sr = mp.data() + items[ITEM_WORDS * i];
            for (int o = items[ITEM_WORDS * i + 1]; o < items[ITEM_WORDS * i + 1] + items[ITEM_WORDS * i + 2]; ++o) {
            const int K = sr[0], n_mul = sr[1], flags = sr[2], sum_size = sr[3], out_off = sr[4];
            const int32_t* ot = mp.data() + sr[5] + o * K;
            const int32_t* st = mp.data() + sr[6];
            for (int64_t b = 0; b < B; ++b) {
                std::vector<int32_t> boff(K);
                for (int k = 0; k < K; ++k) {
                    boff[k] = ot[k];
                    const int ne = sr[SR_FIXED + 2 * k];
                    const int32_t* pairs = mp.data() + sr[SR_FIXED + 2 * k + 1];
                    for (int j = 0; j < ne; ++j) {
                        int32_t s = ev[b * n_ev + pairs[2 * j]];
                        const int32_t card = ev_card[pairs[2 * j]];
                        s = s < 0 ? 0 : (s >= card ? card - 1 : s);
                        boff[k] += s * pairs[2 * j + 1];
                    }
                }
                auto load = [&](int k, int32_t off) { return ((sr[7] >> k) & 1) ? ws[(int64_t)off * ldb + b] : cst[off]; };
                double acc = (flags & FLAG_MAX) ? neg_inf<double>() : 0.0;
                for (int s = 0; s < sum_size; ++s) {
                    double prod = 1.0;
                    for (int k = 0; k < n_mul; ++k) prod *= load(k, boff[k] + st[s * K + k]);
                    if (flags & FLAG_MAX)
                        acc = prod > acc ? prod : acc;
                    else
                        acc += prod;
                }
                if (flags & FLAG_DIV) {
                    double den = 1.0;
                    for (int k = n_mul; k < K; ++k) den *= load(k, boff[k]);
                    const double r = acc / den;
                    acc = (r != r) ? 0.0 : r;
                }
                ws[(int64_t)(out_off + o) * ldb + b] = acc;
            }
            }
        }
    }
    const int32_t* segs = mp.data() + mp[5];
    const int32_t* colmap = mp.data() + mp[6];
    const int out_elems = mp[7];
    for (int g = 0; g < mp[4]; ++g) {
        if (!(segs[4 * g + 3] & SEG_NORMALIZE)) continue;
        for (int64_t b = 0; b < B; ++b) {
            double sum = 0;
            for (int i = 0; i < segs[4 * g + 1]; ++i) sum += ws[(int64_t)(segs[4 * g] + i) * ldb + b];
            for (int i = 0; i < segs[4 * g + 1]; ++i) ws[(int64_t)(segs[4 * g] + i) * ldb + b] /= sum;
        }
    }
    for (int64_t b = 0; b < B; ++b)
        for (int j = 0; j < out_elems; ++j) out[b * out_elems + j] = ws[(int64_t)colmap[j] * ldb + b];
    *n_levels_out = mp[0];
    return 0;
}
