// Plumbing shared by the row-partitioned solvers (dist.cu: Jacobi-PCG, dist_amg.cu: AMG-PCG):
// the dlopen'ed NCCL entry points, the per-rank communicator object and the peer-mapped
// (CUDA IPC) buffers ranks write into directly over NVLink.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "common.cuh"

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
extern NcclApi g_nccl;
int load_nccl();

#define NCCL_TRY(expr)                                                                       \
    do {                                                                                     \
        ncclResult_t _r = (expr);                                                            \
        if (_r != ncclSuccess) {                                                             \
            nodal_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                    \
                            g_nccl.GetErrorString(_r));                                      \
            return NODAL_CUDA_ERROR;                                                         \
        }                                                                                    \
    } while (0)

constexpr int P2P_MAXR = 8;

// One symmetric peer-mapped buffer per rank (same size everywhere): peers store straight into
// it over NVLink (CUDA IPC mappings), so an iteration needs no NCCL call.
struct PeerHeap {
    char* shm = nullptr;                   // this rank's buffer
    size_t bytes = 0;
    std::vector<char*> peer;               // host copy of the mapped base pointers (peer[rank] = shm)
    char** peer_dev = nullptr;             // the same on the device
    unsigned long long* seq = nullptr;     // device: a zero-initialised 256-byte block of counters
};

struct nodal_dist {
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, device = 0;
    bool p2p_disabled = false;
    PeerHeap pcg;                          // dist.cu:     4 KB mailbox + the gathered vector
    PeerHeap amg;                          // dist_amg.cu: mailbox + every [owned | halo] vector of the cycle
};

void peer_heap_release(nodal_dist* d, PeerHeap* h);
// Collective: make sure every rank owns a peer-mapped buffer of at least `bytes` whose first
// `zero_bytes` are zero.  Sets *usable; any failure on any rank disables the path everywhere.
int peer_heap_ensure(nodal_ctx* ctx, nodal_dist* d, PeerHeap* h, size_t bytes, size_t zero_bytes,
                     cudaStream_t st, bool* usable);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
constexpr long long P2P_SPIN_LIMIT = 6000000000ll;   // ~3 s of SM clocks, then give up (no hang)
#endif
