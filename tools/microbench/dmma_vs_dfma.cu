// DMMA (mma.sync.aligned.m8n8k4.f64) vs DFMA throughput on one GPU — settles with a measurement whether the fp64 tensor
// path has any FLOP-rate advantage over the FMA pipe on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_vs_dfma dmma_vs_dfma.cu && ./dmma_vs_dfma
// Each kernel keeps ILP independent accumulator chains per thread and runs `iters` dependent steps on each, all operands
// in registers: the rate measured is the issue/pipe rate, nothing else.
#include <cuda_runtime.h>
#include <cstdio>

template <int ILP>
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_dmma(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9 + i; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA fed from shared memory the way a contraction tile would feed it: A fragment (8x4) and B fragment (4x8) are
// re-loaded from shared memory for every mma (two 8-byte LDS per 256 FMAs per lane-group).
template <int ILP>
__global__ void __launch_bounds__(256) k_dmma_lds(double* out, int iters) {
    __shared__ double sa[64 * 33], sb[64 * 33];
    for (int i = threadIdx.x; i < 64 * 33; i += blockDim.x) { sa[i] = 1.0 + i * 1e-9; sb[i] = 1.0 - i * 1e-9; }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = 0; c1[i] = 0; }
    for (int it = 0; it < iters; ++it) {
        const int k0 = (it & 15) * 4;
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            const double a = sa[(k0 + (lane & 3)) * 33 + (lane >> 2) + 8 * (i & 3)];
            const double b = sb[(k0 + (lane & 3)) * 33 + (lane >> 2) + 8 * (i >> 2)];
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F launch) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int r = 0; r < 5; ++r) launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms / 5;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    const int blocks = sms * 8, threads = 256, iters = 4096;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    constexpr int ILP = 8;
    const float t_fma = time_ms([&] { k_dfma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    const float t_mma = time_ms([&] { k_dmma<ILP><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); });
    const float t_lds = time_ms([&] { k_dmma_lds<ILP><<<blocks, threads>>>(out, iters); });
    const double n_thr = (double)blocks * threads;
    const double fma_flops = 2.0 * n_thr * iters * ILP;                 // one FMA per thread per step
    const double mma_flops = 2.0 * 8 * 8 * 4 * (n_thr / 32) * iters * ILP;  // m8n8k4 per warp per step
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"dmma_lds_fed_tflops\": %.2f, "
           "\"dfma_ms\": %.3f, \"dmma_ms\": %.3f, \"dmma_lds_ms\": %.3f, \"err\": \"%s\"}\n",
           p.name, sms, fma_flops / t_fma / 1e9, mma_flops / t_mma / 1e9, mma_flops / t_lds / 1e9, t_fma, t_mma, t_lds,
           cudaGetErrorString(cudaGetLastError()));
    return 0;
}
