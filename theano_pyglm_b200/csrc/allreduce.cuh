// Peer-memory all-reduce protocol pieces shared by allreduce.cu (stand-alone collective) and llgrad_tc.cu (the fused
// kernel's final reduction with the sum over ranks folded in).  See allreduce.cu for the protocol.
#pragma once
#include "common.cuh"

namespace pyglm {

constexpr int kArMaxWorld = 16;
constexpr int kArMaxBlocks = 192;

struct ArPeers {
    double* recv[kArMaxWorld];      // peer r's receive buffer: [2][world][cap]
    unsigned* flag[kArMaxWorld];    // peer r's flags: [world][kArMaxBlocks]
};

// what a kernel needs to take part in one epoch of the collective
struct ArEpoch {
    ArPeers peers;
    int rank, world;
    int64_t cap;
    unsigned epoch;
};
// start the next epoch on `comm` (host side; every rank must do so in the same order); fails if the communicator is not
// connected or `n` doubles exceed its capacity
int ar_begin_epoch(pyglm_b200_comm* comm, int64_t n, ArEpoch* out);

#ifdef __CUDACC__
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
#endif

}  // namespace pyglm
