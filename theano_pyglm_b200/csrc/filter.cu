// K1: spike-history filtering.
//   X[t][pre*B+b] = sum_{k=1..R} ibasis[k-1][b] * S[t-k][pre]
// Reference: convolve_with_basis, pyglm/utils/basis.py:201-236 (zero row prepended :220,
// 'full'[:T] :232-234), called from LinearBasisImpulses.preprocess_data (impulse.py:114-130).
//
// The reference runs B dense FFT convolutions.  Spike trains are sparse small integers
// (~2% of bins non-zero), so this kernel gathers instead: each block owns a tile of TT output
// bins x PC presynaptic columns, compacts the spikes of the tile (plus its R-bin left halo)
// into per-column lists in shared memory, and every output bin sums ibasis rows over the
// spikes inside its own window.  Accumulation is FP64 in increasing spike-time order (the
// order oracle/convolve_with_basis_direct uses), the result is rounded once to the storage
// type and leaves through a padded shared-memory tile so global stores are fully coalesced.
#include "common.cuh"

namespace pyglm {

constexpr int kFiltThreads = 512;
constexpr int kFiltWarps = kFiltThreads / 32;
constexpr int kFiltSub = 64;        // output bins staged through shared memory at a time

struct FiltSmemLayout {
    size_t off_basis, off_out, off_ent, off_idx, off_s, total;
};

template <typename XT>
static FiltSmemLayout filt_layout(int R, int B, int tt, int pc) {
    FiltSmemLayout L;
    const int rows = tt + R;
    size_t o = 0;
    L.off_basis = o; o += (size_t)R * B * sizeof(double);
    L.off_out = o;   o += (size_t)kFiltSub * (pc * B + 1) * sizeof(XT);
    o = (o + 7) & ~(size_t)7;
    L.off_ent = o;   o += (size_t)rows * pc * sizeof(uint16_t);
    L.off_idx = o;   o += (size_t)(rows + 1) * pc * sizeof(uint16_t);
    L.off_s = o;     o += (size_t)rows * (pc + 4);
    L.total = (o + 15) & ~(size_t)15;
    return L;
}

// BEXACT: B is exactly BMAX (no per-basis predicates in the inner loop)
template <typename XT, int BMAX, bool BEXACT>
__global__ void __launch_bounds__(kFiltThreads)
filter_kernel(const uint8_t* __restrict__ S, int64_t T, int N, int halo,
              const double* __restrict__ ibasis, int R, int B,
              XT* __restrict__ X, int64_t ldx, int tt, int pc, FiltSmemLayout L)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double*   sB   = reinterpret_cast<double*>(smem + L.off_basis);   // [R][B]
    XT*       sOut = reinterpret_cast<XT*>(smem + L.off_out);         // [kFiltSub][pc*B+1]
    uint16_t* sEnt = reinterpret_cast<uint16_t*>(smem + L.off_ent);   // [rows][pc]  tile row of the e-th spike of a column
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(smem + L.off_idx);   // [rows+1][pc] #spikes in tile rows [0,i)
    uint8_t*  sS   = smem + L.off_s;                                  // [rows][pc+4]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rows = tt + R;
    const int64_t t0 = (int64_t)blockIdx.x * tt;      // first output bin of the tile
    const int c0 = blockIdx.y * pc;                   // first presynaptic column of the tile
    const int sstride = pc + 4;
    const int ostride = pc * B + 1;

    // ---- stage the spike tile (tile row i <-> S row halo + t0 - R + i) and the basis.
    // One warp per tile row, lanes across columns, eight rows in flight per warp.
    const int64_t g0 = (int64_t)halo + t0 - R;
    const int64_t gmax = (int64_t)halo + T;
    for (int c = lane; c < pc; c += 32) {
        const bool col_ok = c0 + c < N;
        const uint8_t* src = S + (g0 + warp) * N + c0 + c;            // row `warp` of the tile, this lane's column
        uint8_t* dsts = sS + warp * sstride + c;
        for (int i = warp; i < rows; i += 4 * kFiltWarps) {
            uint8_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t g = g0 + i + u * kFiltWarps;
                v[u] = (col_ok && i + u * kFiltWarps < rows && g >= 0 && g < gmax) ? src[(int64_t)u * kFiltWarps * N] : (uint8_t)0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (i + u * kFiltWarps < rows) dsts[u * kFiltWarps * sstride] = v[u];
            src += (int64_t)4 * kFiltWarps * N;
            dsts += 4 * kFiltWarps * sstride;
        }
    }
    for (int e = tid; e < R * B; e += kFiltThreads) sB[e] = ibasis[e];
    __syncthreads();

    // ---- compact each column's spikes: one warp per column, ballot prefix sums
    for (int c = warp; c < pc; c += kFiltWarps) {
        int count = 0;
        for (int i0 = 0; i0 < rows; i0 += 32) {
            const int i = i0 + lane;
            const uint32_t v = (i < rows) ? sS[i * sstride + c] : 0u;
            const uint32_t m = __ballot_sync(0xffffffffu, v != 0u);
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (i < rows) sIdx[i * pc + c] = (uint16_t)pos;
            if (v) sEnt[pos * pc + c] = (uint16_t)i;
            count += __popc(m);
        }
        if (lane == 0) sIdx[rows * pc + c] = (uint16_t)count;
    }
    __syncthreads();

    const int ncol = min(pc, N - c0);
    const int width = ncol * B;
    for (int sub = 0; sub < tt; sub += kFiltSub) {
        if (t0 + sub >= T) break;
        // ---- gather: a warp task = 32 consecutive output bins of one column
        for (int task = warp; task < (kFiltSub / 32) * ncol; task += kFiltWarps) {
            const int tg = task / ncol, c = task - tg * ncol;
            const int ts = tg * 32 + lane;          // output bin within the sub-tile
            const int tl = sub + ts;                // ... within the tile
            const int i = tl + R;                   // its tile row; window = tile rows [tl, tl+R-1]
            const int e0 = sIdx[(sub + tg * 32) * pc + c];
            const int e1 = sIdx[(sub + tg * 32 + 31 + R) * pc + c];
            double acc[BMAX];
#pragma unroll
            for (int b = 0; b < BMAX; ++b) acc[b] = 0.0;
#pragma unroll 1
            for (int e = e0; e < e1; ++e) {
                const int srow = sEnt[e * pc + c];            // warp-uniform broadcast
                const int k = i - srow;                       // lag
                const uint32_t cnt = sS[srow * sstride + c];  // warp-uniform: spike count of that bin
                if (k >= 1 && k <= R) {
                    const double* row = sB + (size_t)(k - 1) * B;
                    if (cnt == 1u) {                          // the usual case: no multiply
#pragma unroll
                        for (int b = 0; b < BMAX; ++b)
                            if (BEXACT || b < B) acc[b] = __dadd_rn(acc[b], row[b]);
                    } else {
                        const double dc = (double)cnt;
#pragma unroll
                        for (int b = 0; b < BMAX; ++b)
                            if (BEXACT || b < B) acc[b] = __dadd_rn(acc[b], __dmul_rn(dc, row[b]));
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < BMAX; ++b)
                if (BEXACT || b < B) sOut[ts * ostride + c * B + b] = (XT)acc[b];
        }
        __syncthreads();
        // ---- coalesced copy-out of the valid part of the sub-tile
        const int nrow = (int)min((int64_t)kFiltSub, T - (t0 + sub));
        XT* dst = X + (t0 + sub) * ldx + (int64_t)c0 * B;
        {
            const XT* src = sOut + warp * ostride + lane;
            XT* drow = dst + (int64_t)warp * ldx + lane;
            for (int ts = warp; ts < nrow; ts += kFiltWarps) {
#pragma unroll
                for (int j = 0; j < (32 * BMAX + 31) / 32; ++j)              // pc <= 32 columns: at most BMAX chunks of 32
                    if (lane + 32 * j < width) drow[32 * j] = src[32 * j];
                src += kFiltWarps * ostride;
                drow += (int64_t)kFiltWarps * ldx;
            }
        }
        __syncthreads();
    }
}

template <typename XT, int BMAX, bool BEXACT>
static int launch_filter_t(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                           XT* dX, int64_t ldx, cudaStream_t stream)
{
    // pick the largest tile whose shared memory fits (<= 200 KB: one resident block per SM at worst)
    const int cand[][2] = {{192, 32}, {128, 32}, {64, 32}, {64, 16}, {64, 8}, {64, 4}};
    int tt = 0, pc = 0;
    FiltSmemLayout L{};
    for (auto& c : cand) {
        L = filt_layout<XT>(R, B, c[0], c[1]);
        if (L.total <= 113 * 1024) { tt = c[0]; pc = c[1]; break; }     // two resident blocks per SM
    }
    if (tt == 0) {
        set_error("filter: R=%d B=%d needs more shared memory than one SM has", R, B);
        return PYGLM_B200_EUNSUPPORTED;
    }
    if (tt + R > 65535) {
        set_error("filter: R=%d too long", R);
        return PYGLM_B200_EUNSUPPORTED;
    }
    auto kern = filter_kernel<XT, BMAX, BEXACT>;
    PYGLM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total));
    dim3 grid((unsigned)ceil_div(T, tt), (unsigned)ceil_div(N, pc));
    kern<<<grid, kFiltThreads, L.total, stream>>>(dS, T, N, halo, d_ibasis, R, B, dX, ldx, tt, pc, L);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

int launch_filter(const uint8_t* dS, int64_t T, int N, int halo, const double* d_ibasis, int R, int B,
                  void* dX, int64_t ldx, int x_dtype, cudaStream_t stream)
{
    if (B < 1 || B > kMaxBasis) {
        set_error("filter: B=%d outside [1,%d]", B, kMaxBasis);
        return PYGLM_B200_EUNSUPPORTED;
    }
    if (T <= 0) return PYGLM_B200_OK;
    if (x_dtype == PYGLM_B200_X_F32) {
        float* x = static_cast<float*>(dX);
        if (B == 5)  return launch_filter_t<float, 5, true>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        if (B == 10) return launch_filter_t<float, 10, true>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        if (B <= 8)  return launch_filter_t<float, 8, false>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        return launch_filter_t<float, 16, false>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
    } else {
        double* x = static_cast<double*>(dX);
        if (B == 5)  return launch_filter_t<double, 5, true>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        if (B == 10) return launch_filter_t<double, 10, true>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        if (B <= 8)  return launch_filter_t<double, 8, false>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
        return launch_filter_t<double, 16, false>(dS, T, N, halo, d_ibasis, R, B, x, ldx, stream);
    }
}

// ---------------------------------------------------------------------------------
// St[n][t] = S[halo+t][n]: column-major copy of the spikes for the per-column streams
// the Gibbs kernel reads (glm.py:52 indexes S[:, n], a strided column).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transpose_spikes_kernel(const uint8_t* __restrict__ S, int64_t T, int N, int halo, uint8_t* __restrict__ St)
{
    __shared__ uint8_t tile[32][33];
    const int64_t t0 = (int64_t)blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t t = t0 + r;
        const int n = n0 + tx;
        tile[r][tx] = (t < T && n < N) ? S[((int64_t)halo + t) * N + n] : (uint8_t)0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r;
        const int64_t t = t0 + tx;
        if (n < N && t < T) St[(int64_t)n * T + t] = tile[tx][r];
    }
}

int launch_transpose_spikes(const uint8_t* dS, int64_t T, int N, int halo, uint8_t* dSt, cudaStream_t stream)
{
    if (T <= 0) return PYGLM_B200_OK;
    dim3 grid((unsigned)ceil_div(T, 32), (unsigned)ceil_div(N, 32));
    transpose_spikes_kernel<<<grid, 256, 0, stream>>>(dS, T, N, halo, dSt);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

// ---------------------------------------------------------------------------------
// Stimulus features (BasisStimulus, pyglm/components/bkgd.py:122-154): the interpolated stimulus is real
// valued, so its basis projection is a dense causal convolution in FP64,
//   out[t][d*B+b] = sum_{k=1..R} ibasis[k-1][b] * stim[t-k][d]          (utils/basis.py:212-236)
// summed in increasing lag.  One block = 256 bins of one stimulus dimension; window and basis in smem.
// ---------------------------------------------------------------------------------
constexpr int kDenseBins = 256;

__global__ void __launch_bounds__(kDenseBins)
filter_dense_kernel(const double* __restrict__ stim, int64_t T, int D, const double* __restrict__ ibasis, int R, int B,
                    double* __restrict__ out)
{
    extern __shared__ double dsm[];
    double* sIb = dsm;                       // [R][B]
    double* sS = dsm + (size_t)R * B;        // [R + kDenseBins]: stim[t0-R .. t0+255][d]
    const int d = blockIdx.y;
    const int64_t t0 = (int64_t)blockIdx.x * kDenseBins;
    for (int i = threadIdx.x; i < R * B; i += kDenseBins) sIb[i] = ibasis[i];
    for (int i = threadIdx.x; i < R + kDenseBins; i += kDenseBins) {
        const int64_t t = t0 - R + i;
        sS[i] = (t >= 0 && t < T) ? stim[t * D + d] : 0.0;
    }
    __syncthreads();
    const int64_t t = t0 + threadIdx.x;
    if (t >= T) return;
    double acc[kMaxBasis];
#pragma unroll
    for (int b = 0; b < kMaxBasis; ++b) acc[b] = 0.0;
    for (int k = 1; k <= R; ++k) {
        const double v = sS[R + threadIdx.x - k];
#pragma unroll
        for (int b = 0; b < kMaxBasis; ++b)
            if (b < B) acc[b] = fma(sIb[(k - 1) * B + b], v, acc[b]);
    }
#pragma unroll
    for (int b = 0; b < kMaxBasis; ++b)
        if (b < B) out[t * ((int64_t)D * B) + (int64_t)d * B + b] = acc[b];
}

int launch_filter_dense(const double* d_stim, int64_t T, int D, const double* d_ibasis, int R, int B, double* d_out,
                        cudaStream_t stream)
{
    if (T <= 0 || D <= 0) return PYGLM_B200_OK;
    const size_t smem = ((size_t)R * B + R + kDenseBins) * sizeof(double);
    if (smem > 200 * 1024) {
        set_error("filter_dense: R=%d x B=%d basis does not fit in shared memory", R, B);
        return PYGLM_B200_EUNSUPPORTED;
    }
    PYGLM_CUDA(cudaFuncSetAttribute(filter_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(T, kDenseBins), (unsigned)D);
    filter_dense_kernel<<<grid, kDenseBins, smem, stream>>>(d_stim, T, D, d_ibasis, R, B, d_out);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

// X[t][col0 + f] = fstim[t][f]: the stimulus features ride behind the spike-history features of X
template <typename XT>
__global__ void __launch_bounds__(256)
fill_stim_kernel(const double* __restrict__ fstim, int64_t T, int F, XT* __restrict__ X, int64_t ldx, int64_t col0)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= T * F) return;
    const int64_t t = idx / F;
    const int f = (int)(idx - t * F);
    X[t * ldx + col0 + f] = (XT)fstim[idx];
}

int launch_fill_stim(const double* d_fstim, int64_t T, int F, void* dX, int64_t ldx, int64_t col0, int x_dtype,
                     cudaStream_t stream)
{
    if (T <= 0 || F <= 0) return PYGLM_B200_OK;
    const unsigned blocks = (unsigned)ceil_div(T * F, 256);
    if (x_dtype == PYGLM_B200_X_F32) fill_stim_kernel<float><<<blocks, 256, 0, stream>>>(d_fstim, T, F, (float*)dX, ldx, col0);
    else fill_stim_kernel<double><<<blocks, 256, 0, stream>>>(d_fstim, T, F, (double*)dX, ldx, col0);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

}  // namespace pyglm
