// One-shot sum all-reduce of the time shards' partial results over NVLink peer memory.
//
// Time sharding ends every ll+gradient evaluation with a sum over ranks of
// [ll | g_bias | g_w], N (2 + N B) doubles: 29 KB at C2.  The reference performs the same sum on the
// client after its engines return (pyglm/utils/parallel_util.py:30,78; population.py:41-43).  At this
// size a ring or tree collective is pure latency, so every rank instead WRITES its vector into a slot of
// every peer's receive buffer (cudaIpc-mapped device memory, stores travel over NVLink / NVSwitch),
// raises a per-block flag there, waits for the peers' flags and adds the slots up in rank order: one
// kernel, no host round trip, and the same summation order on every rank, so all ranks hold bitwise
// identical results.
//
// Protocol (per block b of the launch, blocks are independent of each other):
//   1. copy slice b of the input into recv[epoch & 1][rank] of every peer (and of this rank)
//   2. __threadfence_system(); flag[rank][b] = epoch at every peer           (release, system scope)
//   3. spin until flag[r][b] >= epoch for every r on this rank                (acquire, system scope)
//   4. out[i] = sum_r recv[epoch & 1][r][i] over slice b, r = 0..world-1, L1-bypassing loads
// Two receive buffers alternate by epoch: a rank can run at most one epoch ahead of a peer, because step 3
// of epoch e needs the peer's flag of epoch e, which the peer raises only after its epoch e-1 kernel ended.
// The grid never exceeds the SM count, so all blocks are resident and step 3 cannot starve step 1.
#include <algorithm>
#include <new>

#include "allreduce.cuh"

namespace pyglm {

constexpr int kArThreads = 256;

__global__ void __launch_bounds__(kArThreads)
allreduce_kernel(ArPeers peers, int rank, int world, int64_t cap, unsigned epoch,
                 const double* __restrict__ in, double* __restrict__ out, int64_t n)
{
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per;
    const int64_t hi = min(n, lo + per);
    const int64_t slot = ((int64_t)(epoch & 1u) * world + rank) * cap;
    for (int r = 0; r < world; ++r) {
        double* dst = peers.recv[r] + slot;
        for (int64_t i = lo + threadIdx.x; i < hi; i += kArThreads) dst[i] = in[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        st_release_sys(peers.flag[threadIdx.x] + (int64_t)rank * kArMaxBlocks + blockIdx.x, epoch);
        const unsigned* mine = peers.flag[rank] + (int64_t)threadIdx.x * kArMaxBlocks + blockIdx.x;
        const long long t0 = clock64();
        // ">= epoch": a peer that already finished this epoch may have raised the flag of the next one
        while ((int)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > (1ll << 37)) __trap();       // a peer never arrived (~70 s): fail loudly, do not hang
        }
    }
    __syncthreads();
    const double* src = peers.recv[rank] + (int64_t)(epoch & 1u) * world * cap;
    for (int64_t i = lo + threadIdx.x; i < hi; i += kArThreads) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcg(src + (int64_t)r * cap + i);
        out[i] = s;
    }
}

}  // namespace pyglm

using namespace pyglm;

struct pyglm_b200_comm {
    int rank = 0, world = 1, device = 0;
    int64_t cap = 0;
    unsigned epoch = 0;
    double* recv = nullptr;       // [2][world][cap]
    unsigned* flag = nullptr;     // [world][kArMaxBlocks]
    ArPeers peers{};
    bool connected = false;
    int num_sms = 0;
};

namespace pyglm {
int ar_begin_epoch(pyglm_b200_comm* c, int64_t n, ArEpoch* out)
{
    PYGLM_REQUIRE(c != nullptr && out != nullptr, "allreduce: null communicator");
    PYGLM_REQUIRE(n >= 0 && n <= c->cap, "allreduce: %lld doubles exceed the communicator capacity %lld", (long long)n, (long long)c->cap);
    if (c->world > 1 && !c->connected) { set_error("allreduce before comm_connect"); return PYGLM_B200_ESTATE; }
    c->epoch += 1;
    if (c->epoch == 0) c->epoch = 1;                       // flags start at 0
    out->peers = c->peers; out->rank = c->rank; out->world = c->world; out->cap = c->cap; out->epoch = c->epoch;
    return PYGLM_B200_OK;
}
}  // namespace pyglm

extern "C" {

int32_t pyglm_b200_comm_world(const pyglm_b200_comm* c) { return c ? c->world : 0; }

int pyglm_b200_comm_create(int32_t rank, int32_t world, int32_t device, int64_t max_doubles, pyglm_b200_comm** out)
{
    PYGLM_REQUIRE(out != nullptr, "comm_create: out is null");
    *out = nullptr;
    PYGLM_REQUIRE(world >= 1 && world <= kArMaxWorld && rank >= 0 && rank < world, "comm_create: bad rank %d / world %d (<= %d)",
                  rank, world, kArMaxWorld);
    PYGLM_REQUIRE(max_doubles >= 1, "comm_create: max_doubles must be positive");
    PYGLM_CUDA(cudaSetDevice(device));
    pyglm_b200_comm* c = new (std::nothrow) pyglm_b200_comm();
    PYGLM_REQUIRE(c != nullptr, "comm_create: host allocation failed");
    c->rank = rank; c->world = world; c->device = device;
    c->cap = round_up(max_doubles, 32);
    cudaError_t e;
    if ((e = cudaMalloc(&c->recv, (size_t)2 * world * c->cap * sizeof(double))) != cudaSuccess ||
        (e = cudaMalloc(&c->flag, (size_t)world * kArMaxBlocks * sizeof(unsigned))) != cudaSuccess ||
        (e = cudaMemset(c->flag, 0, (size_t)world * kArMaxBlocks * sizeof(unsigned))) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess) {
        set_error("comm_create: %s", cudaGetErrorString(e));
        cudaFree(c->recv); cudaFree(c->flag);
        delete c;
        return e == cudaErrorMemoryAllocation ? PYGLM_B200_ENOMEM : PYGLM_B200_ECUDA;
    }
    PYGLM_CUDA(cudaDeviceSynchronize());
    *out = c;
    return PYGLM_B200_OK;
}

int pyglm_b200_comm_export(const pyglm_b200_comm* c, void* handles128)
{
    PYGLM_REQUIRE(c != nullptr && handles128 != nullptr, "comm_export: null argument");
    PYGLM_CUDA(cudaSetDevice(c->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h[2];
    PYGLM_CUDA(cudaIpcGetMemHandle(&h[0], c->recv));
    PYGLM_CUDA(cudaIpcGetMemHandle(&h[1], c->flag));
    memcpy(handles128, h, sizeof(h));
    return PYGLM_B200_OK;
}

int pyglm_b200_comm_connect(pyglm_b200_comm* c, const void* all_handles)
{
    PYGLM_REQUIRE(c != nullptr && all_handles != nullptr, "comm_connect: null argument");
    PYGLM_CUDA(cudaSetDevice(c->device));
    const cudaIpcMemHandle_t* h = static_cast<const cudaIpcMemHandle_t*>(all_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) {
            c->peers.recv[r] = c->recv;
            c->peers.flag[r] = c->flag;
            continue;
        }
        void *pr = nullptr, *pf = nullptr;
        PYGLM_CUDA(cudaIpcOpenMemHandle(&pr, h[2 * r], cudaIpcMemLazyEnablePeerAccess));
        PYGLM_CUDA(cudaIpcOpenMemHandle(&pf, h[2 * r + 1], cudaIpcMemLazyEnablePeerAccess));
        c->peers.recv[r] = static_cast<double*>(pr);
        c->peers.flag[r] = static_cast<unsigned*>(pf);
    }
    c->connected = true;
    return PYGLM_B200_OK;
}

int pyglm_b200_allreduce_sum_dev(pyglm_b200_comm* c, const double* d_in, double* d_out, int64_t n, void* stream)
{
    PYGLM_REQUIRE(c != nullptr, "allreduce: null communicator");
    PYGLM_REQUIRE(n >= 0 && n <= c->cap, "allreduce: %lld doubles exceed the communicator capacity %lld", (long long)n, (long long)c->cap);
    if (n == 0) return PYGLM_B200_OK;
    PYGLM_REQUIRE(d_in && d_out, "allreduce: null buffer");
    PYGLM_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (c->world == 1) {
        if (d_in != d_out) PYGLM_CUDA(cudaMemcpyAsync(d_out, d_in, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        return PYGLM_B200_OK;
    }
    ArEpoch ep;
    { int rc = ar_begin_epoch(c, n, &ep); if (rc) return rc; }
    int blocks = (int)std::min<int64_t>(ceil_div(n, 2 * kArThreads), std::min(kArMaxBlocks, c->num_sms));   // >= 512 doubles per block
    if (blocks < 1) blocks = 1;
    allreduce_kernel<<<blocks, kArThreads, 0, st>>>(c->peers, c->rank, c->world, c->cap, c->epoch, d_in, d_out, n);
    PYGLM_CUDA(cudaGetLastError());
    return PYGLM_B200_OK;
}

int pyglm_b200_comm_destroy(pyglm_b200_comm* c)
{
    if (!c) return PYGLM_B200_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || !c->connected) continue;
        if (c->peers.recv[r]) cudaIpcCloseMemHandle(c->peers.recv[r]);
        if (c->peers.flag[r]) cudaIpcCloseMemHandle(c->peers.flag[r]);
    }
    cudaFree(c->recv);
    cudaFree(c->flag);
    delete c;
    return PYGLM_B200_OK;
}

}  // extern "C"
