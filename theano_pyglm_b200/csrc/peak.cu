// Measured denominators for kernels whose bound is not in MEASURED_PEAKS.json: FP64 FMA throughput of the CUDA cores
// (the bound of K4, the Gibbs delta-ll kernel).  A register-resident DFMA loop, 8 independent chains per thread.
#include "common.cuh"

namespace pyglm {

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* sink, int iters, double a, double b)
{
    double x0 = threadIdx.x * 1e-9, x1 = x0 + 1e-3, x2 = x0 + 2e-3, x3 = x0 + 3e-3, x4 = x0 + 4e-3, x5 = x0 + 5e-3,
           x6 = x0 + 6e-3, x7 = x0 + 7e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;            // never true: keeps the loop alive
}

}  // namespace pyglm

using namespace pyglm;

extern "C" int pyglm_b200_measure_fp64_peak(int32_t device, double* out_tflops)
{
    PYGLM_REQUIRE(out_tflops != nullptr, "measure_fp64_peak: out is null");
    PYGLM_CUDA(cudaSetDevice(device));
    int sms = 0;
    PYGLM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* sink = nullptr;
    PYGLM_CUDA(cudaMalloc(&sink, sizeof(double)));
    cudaEvent_t e0, e1;
    PYGLM_CUDA(cudaEventCreate(&e0));
    PYGLM_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        PYGLM_CUDA(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, 256>>>(sink, iters, 0.999999, 1e-7);
        PYGLM_CUDA(cudaEventRecord(e1));
        PYGLM_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        PYGLM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 64.0 * iters * 256.0 * blocks;
        if (rep > 0 && ms > 0.f) best = best > flops / (ms * 1e-3) / 1e12 ? best : flops / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    *out_tflops = best;
    return PYGLM_B200_OK;
}
