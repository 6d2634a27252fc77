// Microbenchmark: shared-memory read rate of LDS.64 (32 lanes x 8 B consecutive) against LDS.128 and LDS.32,
// in bytes per clock and SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_rate lds_rate.cu && ./lds_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int W>
__global__ void __launch_bounds__(512) k(long long* out, double* sink, int iters)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    const long long t0 = clock64();
    int off = warp * 7;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int base = (off + u * 37) & 2047;
            if (W == 8) {
                a0 += sm[base + lane];
            } else if (W == 16) {
                const double2 v = *reinterpret_cast<const double2*>(&sm[(base & ~1) + 2 * lane]);
                a0 += v.x; a1 += v.y;
            } else {
                a0 += (double)reinterpret_cast<const float*>(sm)[base + lane];
            }
        }
        off += 13;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
}
template <int W> void run(const char* name)
{
    long long* d; double* s; cudaMalloc(&d, 148 * 8 * 8); cudaMalloc(&s, 148 * 2 * 512 * 8);
    const int iters = 2000;
    k<W><<<148 * 2, 512, 4096 * 8 + 64>>>(d, s, iters);
    cudaDeviceSynchronize();
    long long h[296]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double cyc = 0; for (int i = 0; i < 296; ++i) cyc += h[i]; cyc /= 296;
    const double bytes_per_sm = 2.0 * 16 * 32 * W * 8.0 * iters;     // 2 blocks x 16 warps x 32 lanes x W bytes x 8 loads
    printf("%s: %.1f bytes/clk/SM (block time %.0f cycles; includes the dependent FP64 adds)\n", name, bytes_per_sm / cyc, cyc);
}
int main() { run<4>("LDS.32 "); run<8>("LDS.64 "); run<16>("LDS.128"); return 0; }
