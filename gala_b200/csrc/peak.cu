// peak.cu -- FP64 DFMA throughput microbenchmark.  MEASURED_PEAKS.json carries HBM and bf16 peaks
// only; the orbit kernels are bound by the CUDA-core FP64 pipe, so the roofline denominator is
// measured here on the same GPU, in the same process, right before the timed region.
#include <cuda_runtime.h>
#include "../../include/gala_b200.h"

// 8 independent FMA chains per thread keep the FP64 pipe full at any occupancy.
__global__ void __launch_bounds__(256) k_dfma_peak(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3.;
    double x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

extern "C" {
// Returns measured FP64 TFLOP/s (2 flops per DFMA) on the current device, or a negative CUDA error.
double gb_fp64_peak_tflops(int reps) {
    cudaDeviceProp prop; int dev;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return -1.;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    double* out;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)) != cudaSuccess) return -2.;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dfma_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);   // warm-up
    cudaDeviceSynchronize();
    double best = 0.;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        k_dfma_peak<<<blocks, threads>>>(out, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    if (cudaGetLastError() != cudaSuccess) return -3.;
    return best;
}
}
