// dcr_bfc_cuda_edges.cu — cuda-flavour BFC of a symmetric 0/1 adjacency, one work item per UNDIRECTED edge, single- and
// multi-GPU.
//
// Takes over _balanced_forman_curvature / balanced_forman_curvature (curvature/bfc_cuda.py:11-65) at full-graph scale
// (the arxiv-shaped graph: 1.17 M edges).  For such A the kernel's loop collapses exactly (see
// dcr_bfc_cuda.cu): C[i,j] = C[j,i] is a function of d_i, d_j, the support tri(i,j) = A2[i,j] and the numbers of common
// neighbours w with tri(i,w) = 1 resp. tri(j,w) = 1.  The per-entry kernels of dcr_bfc_cuda.cu compute every edge
// twice (once per direction) and find the row of an entry by binary search; here
//   * pass 1 computes tri(e) once per undirected edge (sorted-list intersection: lanes take the shorter row and
//     binary-search the longer one),
//   * pass 2 repeats the intersection, reads the two supports at the matched positions through `eid` (directed entry
//     -> undirected edge id, built once per graph), and evaluates the closing formula once,
//   * edges whose shorter row has >= HEAVY_MIN entries (hub-hub edges; a few thousand of a million, but each would
//     hold one warp for the length of the whole pass) are taken FIRST, one CTA per edge, by the same persistent launch.
// Multi-GPU: contiguous edge ranges, graph replicated; pass 2 needs the supports of OTHER ranks'
// edges, so pass 1 stores each support into every rank's buffer (dcr_comm: CUDA-IPC peer memory over NVLink — the
// all-gather of `tri` is fused into the kernel), the ranks meet on device-side flags, and pass 2 delivers the results
// the same way.  Results per edge in the dcr_comm layout: c64 f64[chunk] | tri | sharp | lam | c32 (fp32 bits).
// Memory-system bound (gathers served by L2); algorithmic bytes per edge: 16 + 4(d_i+d_j) + 8 tri + 24.
#include <algorithm>

#include "dcr_comm.cuh"

namespace dcr {

constexpr int HEAVY_MIN = 512;     // shorter row >= this: one CTA per edge
constexpr int CE_THREADS = 256;

struct EdgeArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const int32_t* esrc;
    const int32_t* edst;
    const int32_t* eid;        // [nnz] directed entry -> undirected edge id
    const int32_t* heavy;      // heavy[0] = count, heavy[8..] = edge ids
    int64_t lo, hi;            // this call's edge range
    // local outputs, indexed by edge id
    int32_t* tri; int32_t* sharp; int32_t* lam; double* c64; float* c32;
};

__device__ __forceinline__ bool is_heavy(int di, int dj) { return min(di, dj) >= HEAVY_MIN; }

// store one 32-bit result at the edge's position of array `which` (0 = tri, 1 = sharp, 2 = lam, 3 = c32) in every peer
__device__ __forceinline__ void peers_store32(const CommView& c, int which, int64_t e, int32_t v) {
    for (int q = 1; q < c.world; ++q) {
        const int p = (c.rank + q) % c.world;
        ((int32_t*)(c.peers[p] + c.chunk * 8))[which * c.chunk + e] = v;
    }
}
__device__ __forceinline__ void peers_store64(const CommView& c, int64_t e, double v) {
    for (int q = 1; q < c.world; ++q) ((double*)c.peers[(c.rank + q) % c.world])[e] = v;
}

// ---- set-up, once per graph ------------------------------------------------------------------------------------
__global__ void edges_prepare_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                     const int32_t* __restrict__ esrc, const int32_t* __restrict__ edst, int64_t n_edges,
                                     int32_t* __restrict__ eid, int32_t* __restrict__ heavy) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int i = esrc[e], j = edst[e];
    const int si = rowptr[i], di = rowptr[i + 1] - si, sj = rowptr[j], dj = rowptr[j + 1] - sj;
    const int f = find_sorted(colidx, si, di, j), r = find_sorted(colidx, sj, dj, i);
    if (f >= 0) eid[f] = (int32_t)e;
    if (r >= 0) eid[r] = (int32_t)e;
    if (is_heavy(di, dj)) heavy[8 + atomicAdd(&heavy[0], 1)] = (int32_t)e;
}

// ---- pass 1: supports ------------------------------------------------------------------------------------------
// `units` of THREADS_PER_EDGE threads walk the shorter row; every thread of the unit returns the count
template <int HEAVY>
__device__ __forceinline__ int edge_support(const EdgeArgs& a, int i, int j, int* red) {
    int si = a.rowptr[i], di = a.rowptr[i + 1] - si, sj = a.rowptr[j], dj = a.rowptr[j + 1] - sj;
    if (di > dj) { int t = di; di = dj; dj = t; t = si; si = sj; sj = t; }
    int c = 0;
    if (HEAVY) {
        for (int t = threadIdx.x; t < di; t += CE_THREADS) c += find_sorted(a.colidx, sj, dj, a.colidx[si + t]) >= 0;
        c = warp_sum(c);
        __syncthreads();
        if (threadIdx.x == 0) red[0] = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[0], c);
        __syncthreads();
        return red[0];
    }
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < di; t += 32) c += find_sorted(a.colidx, sj, dj, a.colidx[si + t]) >= 0;
    return warp_sum(c);
}

// One persistent launch per pass: every CTA first takes hub-hub edges (one CTA per edge, block-strided over the list
// built by the set-up kernel), then its warps take the ordinary edges — the hub edges start first and no second launch
// waits for the tail of the first.
__global__ void __launch_bounds__(CE_THREADS) edges_support_kernel(EdgeArgs a, CommView c) {
    __shared__ int red[2];
    __shared__ int s_last;
    if (c.world > 1) {                    // the peers' buffers are free to take this pass
        CommFlags* fl = comm_flags(c.peers[c.rank], c.flag_off);
        if (threadIdx.x < c.world && threadIdx.x != c.rank) comm_spin(fl, &fl->ready[threadIdx.x], c.epoch);
        __syncthreads();
    }
    const int nh = a.heavy[0];
    for (int h = blockIdx.x; h < nh; h += gridDim.x) {
        const int64_t e = a.heavy[8 + h];
        if (e < a.lo || e >= a.hi) continue;
        const int t = edge_support<1>(a, a.esrc[e], a.edst[e], red);
        if (threadIdx.x == 0) { a.tri[e] = t; peers_store32(c, 0, e, t); }
    }
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = a.lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < a.hi; e += warps) {
        const int i = a.esrc[e], j = a.edst[e];
        if (is_heavy(a.rowptr[i + 1] - a.rowptr[i], a.rowptr[j + 1] - a.rowptr[j])) continue;
        const int t = edge_support<0>(a, i, j, red);
        if (lane == 0) { a.tri[e] = t; peers_store32(c, 0, e, t); }
    }
    comm_block_done(c, &s_last, c.epoch);
}

// ---- pass 2: "support == 1" counts at the common neighbours + closing formula ----------------------------------
template <int HEAVY>
__device__ __forceinline__ void edge_closing(const EdgeArgs& a, const CommView& c, int64_t e, int* red) {
    const int i = a.esrc[e], j = a.edst[e];
    int si = a.rowptr[i], di = a.rowptr[i + 1] - si, sj = a.rowptr[j], dj = a.rowptr[j + 1] - sj;
    const int dmax = max(di, dj), dmin = min(di, dj), dsum = di + dj;
    if (di > dj) { int t = di; di = dj; dj = t; t = si; si = sj; sj = t; }
    int ones = 0;       // #{w common : tri(short side, w) == 1} + #{w common : tri(long side, w) == 1}
    const int step = HEAVY ? CE_THREADS : 32, first = HEAVY ? threadIdx.x : (threadIdx.x & 31);
    for (int t = first; t < di; t += step) {
        const int q = find_sorted(a.colidx, sj, dj, a.colidx[si + t]);
        if (q >= 0) ones += (a.tri[a.eid[si + t]] == 1) + (a.tri[a.eid[q]] == 1);
    }
    ones = warp_sum(ones);
    if (HEAVY) {
        __syncthreads();
        if (threadIdx.x == 0) red[0] = 0;
        __syncthreads();
        if ((threadIdx.x & 31) == 0) atomicAdd(&red[0], ones);
        __syncthreads();
        ones = red[0];
    }
    if ((HEAVY ? threadIdx.x : (threadIdx.x & 31)) == 0) {
        // bfc_cuda.py:20-48 for symmetric 0/1 A: sharp = d_i + d_j - t1 - t2, lambda = d_max (App. A.2)
        const int sharp = dsum - ones, lam = dmax;
        const Closing v = closing_value(dmax, dmin, a.tri[e], 1, sharp, lam);
        if (a.sharp) a.sharp[e] = sharp;
        if (a.lam) a.lam[e] = lam;
        if (a.c64) a.c64[e] = v.c64;
        a.c32[e] = v.c32;
        if (c.world > 1) {
            peers_store32(c, 1, e, sharp);
            peers_store32(c, 2, e, lam);
            peers_store32(c, 3, e, __float_as_int(v.c32));
            peers_store64(c, e, v.c64);
        }
    }
}

__global__ void __launch_bounds__(CE_THREADS) edges_closing_kernel(EdgeArgs a, CommView c, unsigned int wait_epoch) {
    __shared__ int red[2];
    __shared__ int s_last;
    if (c.world > 1) {                    // every rank's supports of this pass have landed in the local buffer
        CommFlags* fl = comm_flags(c.peers[c.rank], c.flag_off);
        if (threadIdx.x < c.world && threadIdx.x != c.rank) comm_spin(fl, &fl->done[threadIdx.x], wait_epoch);
        __syncthreads();
    }
    const int nh = a.heavy[0];
    for (int h = blockIdx.x; h < nh; h += gridDim.x) {
        const int64_t e = a.heavy[8 + h];
        if (e < a.lo || e >= a.hi) continue;
        edge_closing<1>(a, c, e, red);
    }
    __syncthreads();
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = a.lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); e < a.hi; e += warps) {
        const int i = a.esrc[e], j = a.edst[e];
        if (is_heavy(a.rowptr[i + 1] - a.rowptr[i], a.rowptr[j + 1] - a.rowptr[j])) continue;
        edge_closing<0>(a, c, e, red);
    }
    comm_block_done(c, &s_last, c.epoch);
}

}  // namespace dcr

using namespace dcr;

extern "C" int64_t dcr_bfc_cuda_edges_aux_ints(int64_t nnz, int64_t n_edges) { return nnz + 8 + n_edges; }

extern "C" int dcr_bfc_cuda_edges_prepare(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* esrc,
                                          const int32_t* edst, int64_t n_edges, int64_t nnz, int32_t* aux, void* stream) {
    (void)n;
    if (n_edges <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DCR_CUDA(cudaMemsetAsync(aux + nnz, 0, 8 * sizeof(int32_t), st));
    edges_prepare_kernel<<<(unsigned)((n_edges + 255) / 256), 256, 0, st>>>(rowptr, colidx, esrc, edst, n_edges, aux,
                                                                          aux + nnz);
    DCR_LAUNCH_CHECK();
    return 0;
}

static int edges_grid(int64_t count) {
    const int64_t cap = (int64_t)sm_count() * 8;
    return (int)std::max<int64_t>(1, std::min<int64_t>((count + 7) / 8, cap));
}

// phases: bit 0 = supports, bit 1 = closing.  c == nullptr: single GPU.
static int edges_run(const int32_t* rowptr, const int32_t* colidx, const int32_t* esrc, const int32_t* edst, int64_t nnz,
                     const int32_t* aux, int64_t e_lo, int64_t count, int32_t* tri, int32_t* sharp, int32_t* lam,
                     double* c64, float* c32, int phases, dcr_comm* c, cudaStream_t st) {
    EdgeArgs a;
    a.rowptr = rowptr; a.colidx = colidx; a.esrc = esrc; a.edst = edst; a.eid = aux; a.heavy = aux + nnz;
    a.lo = e_lo; a.hi = e_lo + count;
    a.tri = tri; a.sharp = sharp; a.lam = lam; a.c64 = c64; a.c32 = c32;
    CommView v;
    v.peers = nullptr; v.rank = 0; v.world = 1; v.chunk = 0; v.flag_off = 0; v.epoch = 0;
    unsigned int epoch_support = 0;
    const int grid = edges_grid(count);
    if (phases & 1) {
        if (c) { v = comm_view(c, ++c->epoch); epoch_support = v.epoch; }
        if (c && c->world > 1) { comm_ready_kernel<<<1, COMM_MAX_WORLD, 0, st>>>(v); DCR_LAUNCH_CHECK(); }
        edges_support_kernel<<<grid, CE_THREADS, 0, st>>>(a, v);
        DCR_LAUNCH_CHECK();
    }
    if (phases & 2) {
        if (c) v = comm_view(c, ++c->epoch);
        edges_closing_kernel<<<grid, CE_THREADS, 0, st>>>(a, v, epoch_support);
        DCR_LAUNCH_CHECK();
        if (c && c->world > 1) { comm_wait_kernel<<<1, COMM_MAX_WORLD, 0, st>>>(v); DCR_LAUNCH_CHECK(); }
    }
    return 0;
}

extern "C" int dcr_bfc_cuda_edges(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* esrc,
                                  const int32_t* edst, int64_t nnz, const int32_t* aux, int64_t e_lo, int64_t count,
                                  int32_t* tri, int32_t* sharp, int32_t* lam, double* c64, float* c32, int phases,
                                  void* stream) {
    (void)n;
    if (count <= 0) return 0;
    if (!tri || ((phases & 2) && !c32) || !aux) { set_error("dcr_bfc_cuda_edges: tri, c32 and aux must not be NULL"); return 1; }
    return edges_run(rowptr, colidx, esrc, edst, nnz, aux, e_lo, count, tri, sharp, lam, c64, c32, phases, nullptr,
                     (cudaStream_t)stream);
}

extern "C" int dcr_bfc_cuda_sharded(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* esrc,
                                    const int32_t* edst, int64_t nnz, const int32_t* aux, int64_t e_lo, int64_t count,
                                    dcr_comm* c, void* stream) {
    (void)n;
    if (!c || !aux) { set_error("dcr_bfc_cuda_sharded: comm / aux is NULL"); return 1; }
    if (!c->connected) { set_error("dcr_bfc_cuda_sharded: dcr_comm_connect has not been called"); return 1; }
    if (e_lo < 0 || count < 0 || e_lo + count > c->n_edges) { set_error("dcr_bfc_cuda_sharded: edge range outside [0, n_edges)"); return 1; }
    int32_t* ints = (int32_t*)(c->local + c->chunk * 8);
    return edges_run(rowptr, colidx, esrc, edst, nnz, aux, e_lo, count, ints, ints + c->chunk, ints + 2 * c->chunk,
                     (double*)c->local, (float*)(ints + 3 * c->chunk), 3, c, (cudaStream_t)stream);
}
