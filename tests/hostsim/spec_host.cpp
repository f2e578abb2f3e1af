// Test infrastructure only: runs the source pgx_spec.cu GENERATES for a plan (straight-line CUDA C) on the CPU, one
// thread at a time and barrier phase by barrier phase, so the generator's indexing, lifetime packing and constant folding are checked against the numpy plan
// interpreter without a GPU. Compile:  g++ -O1 -shared -fPIC -DPGX_GENERATED='"file.cu"' spec_host.cpp
#include <cmath>
#include <cstdint>
#include <cstring>

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static Dim3 threadIdx, blockIdx, gridDim;
alignas(16) unsigned char smem_raw[256 * 1024];
#define __global__
#define __device__ static
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__
#define PGX_HOST_SIM 1
#define __noinline__
#define __syncwarp()
#define __syncthreads()
static inline double __ldg(const double* p) { return *p; }
static inline float __ldg(const float* p) { return *p; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline double __longlong_as_double(long long v) { double d; memcpy(&d, &v, 8); return d; }
static int pgx_phase_v = 0;
#define PGX_PHASE pgx_phase_v

#include PGX_GENERATED

extern "C" void spec_host_run(const void* cst, const int* ev, void* out, long long B) {
    gridDim.x = (unsigned)((B + 32 * PGX_ROWS - 1) / (32 * PGX_ROWS));  // one row per block: a persistent kernel loops once
    for (long long blk = 0; blk * 32 * PGX_ROWS < B; ++blk) {
        blockIdx.x = (unsigned)blk;
        for (int phase = 0; phase < PGX_N_PHASES; ++phase) {
            pgx_phase_v = phase;
            for (unsigned t = 0; t < 32u * PGX_WARPS * PGX_ROWS; ++t) {
                threadIdx.x = t;
                k_plan_spec((const T*)cst, ev, (T*)out, B);
            }
        }
    }
}
