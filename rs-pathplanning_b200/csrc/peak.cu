// peak.cu -- FP64 pipe peak micro-benchmark: 8 independent DFMA chains per
// thread so that the pipe, not the dependency latency, is the limit.
#include "pp_common.cuh"

#define PP_PEAK_THREADS 256
#define PP_PEAK_CHAINS 8
#define PP_PEAK_INNER 64

__global__ void __launch_bounds__(PP_PEAK_THREADS) pp_fp64_peak_kernel(int iters, double *sink) {
    double a[PP_PEAK_CHAINS];
    const double m = 1.0000000001, c = 1e-9 * (double)(threadIdx.x + 1);
#pragma unroll
    for (int k = 0; k < PP_PEAK_CHAINS; ++k) a[k] = (double)(k + 1) + (double)blockIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < PP_PEAK_INNER; ++u) {
#pragma unroll
            for (int k = 0; k < PP_PEAK_CHAINS; ++k) a[k] = __fma_rn(a[k], m, c);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < PP_PEAK_CHAINS; ++k) s += a[k];
    if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

int pp_launch_fp64_peak(pp_ctx *ctx, int iters, double *sink, cudaStream_t stream, unsigned *threads,
                        unsigned *per_thread_per_iter) {
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    pp_launch_scope scope(ctx, "fp64_peak");
    pp_fp64_peak_kernel<<<grid, PP_PEAK_THREADS, 0, stream>>>(iters, sink);
    PP_CUDA(ctx, cudaGetLastError());
    *threads = grid * PP_PEAK_THREADS;
    *per_thread_per_iter = PP_PEAK_CHAINS * PP_PEAK_INNER;
    return PP_OK;
}
