// dcr_comm.cuh — peer-memory exchange shared by the multi-GPU kernels (paper flavour: dcr_bfc_paper.cu, cuda flavour:
// dcr_bfc_cuda_edges.cu).  Every rank owns the FULL per-edge result arrays (24 bytes per undirected edge: one f64 array
// and four 32-bit arrays of `chunk` entries) in a cudaMalloc'ed buffer its peers map through CUDA IPC (NVLink / NVSwitch
// peer memory).  A computing kernel stores each result at the edge's position in EVERY rank's buffer, so compute and
// all-gather are one kernel.  Two small flag arrays per buffer carry the hand-shake: ready[p] = "rank p has passed the
// start of pass k" (its consumers of the previous pass are done: its buffer may be overwritten), done[p] = "all of
// rank p's results of (sub-)pass k have landed".  Epochs only grow; comparisons are wrap-safe.
#pragma once

#include <stddef.h>

#include "dcr_common.cuh"

constexpr int COMM_MAX_WORLD = 32;
struct CommFlags {
    unsigned int ready[COMM_MAX_WORLD];
    unsigned int done[COMM_MAX_WORLD];
    unsigned int blocks_done;        // last-block detection of the closing kernel
    unsigned int error;              // a wait timed out (the peers never arrived)
};
struct dcr_comm {
    int rank, world, device;
    int64_t n_edges, chunk;          // arrays are `chunk` entries long (n_edges rounded up to 4)
    size_t bytes, flag_off;
    unsigned char* local;
    unsigned char* peer[COMM_MAX_WORLD];
    unsigned char** d_peers;         // the same pointers on the device
    unsigned int epoch;
    bool connected;
};

namespace dcr {

struct CommView {
    unsigned char* const* peers;     // [world] buffers
    int rank, world;
    int64_t chunk;
    size_t flag_off;
    unsigned int epoch;
};
__device__ __forceinline__ CommFlags* comm_flags(unsigned char* buf, size_t flag_off) { return (CommFlags*)(buf + flag_off); }
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long comm_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr unsigned long long COMM_TIMEOUT_NS = 4000000000ull;     // a peer that has not arrived after 4 s never will

// start of a pass: tell every peer that this rank's buffer may be overwritten
static __global__ void comm_ready_kernel(CommView c) {
    const int p = threadIdx.x;
    if (p < c.world && p != c.rank) st_release_sys(&comm_flags(c.peers[p], c.flag_off)->ready[c.rank], c.epoch);
}

// end of a pass: the peers' results have landed in this rank's buffer
static __global__ void comm_wait_kernel(CommView c) {
    CommFlags* fl = comm_flags(c.peers[c.rank], c.flag_off);
    const int p = threadIdx.x;
    if (p < c.world && p != c.rank) {
        const unsigned long long t0 = comm_now();
        while ((int)(ld_acquire_sys(&fl->done[p]) - c.epoch) < 0) {
            if (comm_now() - t0 > COMM_TIMEOUT_NS) { fl->error = 1u; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}


// ---- pieces a computing kernel is assembled from -------------------------------------------------------------
// threads < world wait until peer `threadIdx.x` has published flags[...] >= epoch in THIS rank's buffer
__device__ __forceinline__ void comm_spin(CommFlags* fl, const unsigned int* flag, unsigned int epoch) {
    const unsigned long long t0 = comm_now();
    while ((int)(ld_acquire_sys(flag) - epoch) < 0) {
        if (comm_now() - t0 > COMM_TIMEOUT_NS) { fl->error = 1u; break; }
        __nanosleep(200);
    }
}
// call by ALL threads of every block after the block's last peer store: the last block to arrive publishes
// done[rank] = epoch in every peer's buffer.  `s_last` is a shared int.
__device__ __forceinline__ void comm_block_done(const CommView& c, int* s_last, unsigned int epoch) {
    if (c.world == 1) return;
    CommFlags* fl = comm_flags(c.peers[c.rank], c.flag_off);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *s_last = (atomicAdd(&fl->blocks_done, 1u) == gridDim.x - 1);
    __syncthreads();
    if (*s_last) {
        __threadfence_system();
        if (threadIdx.x == 0) fl->blocks_done = 0u;
        if (threadIdx.x < c.world && threadIdx.x != c.rank)
            st_release_sys(&comm_flags(c.peers[threadIdx.x], c.flag_off)->done[c.rank], epoch);
    }
}

}  // namespace dcr

static inline dcr::CommView comm_view(const dcr_comm* c, unsigned int epoch) {
    dcr::CommView v;
    v.peers = c->d_peers; v.rank = c->rank; v.world = c->world; v.chunk = c->chunk; v.flag_off = c->flag_off;
    v.epoch = epoch;
    return v;
}
