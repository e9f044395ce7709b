// issue_mix_probe.cu -- does a non-FP64 instruction issue "for free" beside a DFMA on sm_100?
// The FP64 pipe has 16 lanes per SM sub-partition, so a warp-wide DFMA occupies it for 2 cycles.  The probe
// runs 8 independent DFMA chains per thread and adds K independent 32-bit integer (or FSEL-like) operations
// per DFMA, K = 0..3, and prints the time per DFMA: flat up to K = 1 means the scheduler dual-issues around the
// FP64 pipe, a slope means every instruction costs an issue slot on top of the pipe time.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build_ab/issue_mix_probe tools/issue_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K, int WARPS_PER_SMSP>
__global__ void __launch_bounds__(WARPS_PER_SMSP * 128) probe(double *out, unsigned *iout, double a, double b, unsigned s, int iters) {
    double x[8];
    unsigned u[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        x[i] = a + i + threadIdx.x;
        u[i] = s + i * 7 + threadIdx.x;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x[i]) : "d"(a), "d"(b));
#pragma unroll
                for (int k = 0; k < K; ++k)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(u[(i + k) & 7]) : "r"(s), "r"(u[(i + k + 3) & 7]));
            }
        }
    }
    double acc = 0;
    unsigned ua = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc += x[i];
        ua ^= u[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = ua;
}

template <int K, int W>
static void run(double *out, unsigned *iout, int sms) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<K, W><<<sms, W * 128>>>(out, iout, 0.999999, 1e-9, 12345u, 10);
    cudaEventRecord(e0);
    probe<K, W><<<sms, W * 128>>>(out, iout, 0.999999, 1e-9, 12345u, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double dfma_per_smsp = (double)iters * 32 * W;  // warp-level DFMAs each sub-partition issues
    printf("K=%d warps/SMSP=%d  %.3f ms  %.3f ns per warp-DFMA per SMSP  (%s)\n", K, W, ms, ms * 1e6 / dfma_per_smsp,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    unsigned *iout;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    cudaMalloc(&iout, sizeof(unsigned) * p.multiProcessorCount * 1024);
    printf("%s, %d SMs, %.0f MHz nominal max\n", p.name, p.multiProcessorCount, p.clockRate / 1e3);
    run<0, 1>(out, iout, p.multiProcessorCount);
    run<0, 4>(out, iout, p.multiProcessorCount);
    run<1, 4>(out, iout, p.multiProcessorCount);
    run<2, 4>(out, iout, p.multiProcessorCount);
    run<3, 4>(out, iout, p.multiProcessorCount);
    run<0, 6>(out, iout, p.multiProcessorCount);
    run<1, 6>(out, iout, p.multiProcessorCount);
    run<2, 6>(out, iout, p.multiProcessorCount);
    run<3, 6>(out, iout, p.multiProcessorCount);
    return 0;
}
