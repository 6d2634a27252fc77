// K2 (tensor-core path): tcgen05 contractions on FP16 split planes, fused FP32 epilogue, FP64 sums.
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"

namespace pyglm {

// Device-resident derived data the tensor-core path keeps per dataset.
struct TcWorkspace {
    // split planes of the filtered spike train: X * sx = X1 + X2 * 2^-11  (FP16 each)
    __half* X1 = nullptr;
    __half* X2 = nullptr;
    int64_t ldp = 0;              // row pitch of the planes in elements (multiple of 8)
    float* sx = nullptr;          // [NB] per-feature power-of-two scale
    unsigned* colmax = nullptr;   // [NB + N] scratch for the scale search (per feature, per presynaptic column)
    uint8_t* Sp = nullptr;        // [T][Np] zero-padded copy of the spikes (Np % 32 == 0)
    int Np = 0;
    unsigned* colflag = nullptr;  // [N] range flags raised by the FP32 epilogues (exp nonlinearity, see kExpSafe)
    bool planes_ready = false;
    // from-spikes evaluation (spikes-only datasets): X1 / X2 hold one time chunk, refilled by K1 in every evaluation
    bool streamed = false;
    int64_t chunk_rows = 0;       // rows of the chunk buffers
    int64_t map_rows = 0;         // rows the tensor maps over X1 / X2 were encoded for
    double* acc = nullptr;        // per-chunk ll / g_bias / g_w before they are added into the outputs
    size_t acc_elems = 0;
    // per-call operands
    __half* Mp = nullptr;         // [2][32][Kp] split planes of the scaled weight matrix
    float* colpar = nullptr;      // [2][32]: 1/sm[n], bias[n]
    double* part = nullptr;       // per-CTA partial sums
    size_t part_elems = 0;
    void* tmaps = nullptr;        // host copy of the two CUtensorMap descriptors
    int num_sms = 0;
    void* gemm = nullptr;         // large-population GEMM path workspace (llgrad_tc_gemm.cu)
    void release();
};

struct TcArgs {
    const float* X; int64_t ldx;
    const uint8_t* S; int64_t T; int N; int halo; int B; int F;
    const double* ibasis; int R;  // [R][B] interpolated basis (scales of the split planes)
    double dt; int nlin;
    int n_lo, ncols;
    const double* bias; const double* w; const int8_t* A; const double* W;
    double* out_ll; double* out_gb; double* out_gw;
    unsigned* flags;              // [N] range flags (index = neuron) or nullptr
    const void* ar = nullptr;     // ArEpoch*: fold the sum over ranks into the final reduction (out_ll = the whole result vector)
};

bool tc_supported(int64_t T, int N, int B, int x_dtype);
bool tc_uses_fused_kernel(int64_t nfeat);
bool tc_can_fuse_allreduce(int N, int64_t nfeat);
// shared with the GEMM path
int tc_ensure_planes(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream);
int tc_build_planes_direct(TcWorkspace& ws, const uint8_t* S, int64_t T, int N, int halo, const double* d_ibasis,
                           int R, int B, float* X, int64_t ldx, cudaStream_t stream);
int tc_prepare_streamed(TcWorkspace& ws, const uint8_t* S, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                        cudaStream_t stream);
int launch_tc_ll_grad_streamed(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream);
int tc_make_map_2d(void* map, const void* base, int64_t dim0, int64_t dim1, int64_t pitch_elems, int box0, int box1,
                   int swizzle = 0);        // 0: SWIZZLE_64B, 1: SWIZZLE_128B, 2: SWIZZLE_32B
int launch_tc_gemm_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream);
void tc_gemm_release(void* w);
int launch_tc_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream);

}  // namespace pyglm
