// K2 (tensor-core path): tcgen05 3xTF32 contractions with fused FP32 epilogue and FP64 sums.
#pragma once
#include "common.cuh"

namespace pyglm {

struct TcWorkspace {
    void* buf = nullptr;      // device scratch (operand splits, partial sums)
    size_t bytes = 0;
    void* tmap = nullptr;     // cached tensor maps (host)
    void release();
};

struct TcArgs {
    const float* X; int64_t ldx;
    const uint8_t* S; int64_t T; int N; int halo; int B;
    double dt; int nlin;
    int n_lo, ncols;
    const double* bias; const double* w; const int8_t* A; const double* W;
    double* out_ll; double* out_gb; double* out_gw;
};

bool tc_supported(int64_t T, int N, int B, int x_dtype);
int launch_tc_ll_grad(const TcArgs& a, TcWorkspace& ws, cudaStream_t stream);

}  // namespace pyglm
