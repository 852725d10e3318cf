// dcr_bfc_cuda.cu — cuda-flavour Balanced Forman curvature over a sorted CSR.
//
// Takes over _balanced_forman_curvature (curvature/bfc_cuda.py:11-48) + its host wrapper (:51-65) for a symmetric
// 0/1 adjacency without self-loops.  The reference spends O(N^3): a dense A@A (:53) and an N^2-thread kernel with
// an N-long loop (:33).  For such A the loop collapses exactly:
//     k == i  -> T1 = d_i,  k == j -> T2 = d_j                      (the endpoints count themselves)
//     k in N(j)\{i}: T1 = c(i,k) - A[i,k] >= 0, zero iff k is a triangle node with support(i,k) == 1
//     k in N(i)\{j}: T2 = c(k,j) - A[k,j],      zero iff k is a triangle node with support(k,j) == 1
//   => sharp = d_i + d_j - t1 - t2,  lambda = max(d_i, d_j),  A2[i,j] = support(i,j)
// so the work is two passes of sorted-list intersections: supports of all entries, then per entry the two
// "support == 1" counts read at the matched positions.  One warp per directed entry; lanes take elements of the
// shorter row (coalesced) and binary-search the longer one.  Memory-system bound (gathers served by L2);
// algorithmic bytes per undirected edge: 16 + 4(d_i+d_j) + 8 tri + 24.
#include "dcr_directed.cuh"

namespace dcr {

// row of directed entry p: largest v with rowptr[v] <= p  (rows may be empty)
__device__ __forceinline__ int row_of_entry(const int32_t* __restrict__ rowptr, int n, int64_t p) {
    int a = 0, b = n;  // invariant: rowptr[a] <= p < rowptr[b]
    while (b - a > 1) {
        int m = (a + b) >> 1;
        if ((int64_t)rowptr[m] <= p) a = m; else b = m;
    }
    return a;
}

__global__ void __launch_bounds__(256) support_kernel(const int32_t* __restrict__ rowptr,
                                                      const int32_t* __restrict__ colidx, int n,
                                                      int32_t* __restrict__ tri, int64_t lo, int64_t hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    GraphView g{rowptr, nullptr, colidx};
    for (int64_t p = lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < hi; p += warps) {
        const int i = row_of_entry(rowptr, n, p);
        const int j = colidx[p];
        const int c = warp_intersect_count(g, i, j, lane);
        if (lane == 0) tri[p] = c;
    }
}

__global__ void __launch_bounds__(256) cuda_flavour_kernel(const int32_t* __restrict__ rowptr,
                                                           const int32_t* __restrict__ colidx, int n,
                                                           const int32_t* __restrict__ tri,
                                                           int32_t* __restrict__ sharp_out,
                                                           int32_t* __restrict__ lam_out, double* __restrict__ c64_out,
                                                           float* __restrict__ c32_out, int64_t lo, int64_t hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < hi; p += warps) {
        const int i = row_of_entry(rowptr, n, p);
        const int j = colidx[p];
        int si = rowptr[i], di = rowptr[i + 1] - si;
        int sj = rowptr[j], dj = rowptr[j + 1] - sj;
        // walk the shorter row, search the longer; remember which side is "i" for t1/t2
        const bool swap = di > dj;
        const int sa = swap ? sj : si, da = swap ? dj : di;
        const int sb = swap ? si : sj, db = swap ? di : dj;
        int ta = 0, tb = 0;  // #triangle nodes w with support(a-side,w)==1 / support(b-side,w)==1
        for (int t = lane; t < da; t += 32) {
            const int w = colidx[sa + t];
            const int q = find_sorted(colidx, sb, db, w);
            if (q >= 0) {
                ta += (tri[sa + t] == 1);
                tb += (tri[q] == 1);
            }
        }
        ta = warp_sum(ta);
        tb = warp_sum(tb);
        if (lane == 0) {
            // bfc_cuda.py:20-29: d_max/d_min from d_in[i], d_out[j]; an entry exists so both degrees are >= 1
            const int dmax = max(di, dj), dmin = min(di, dj);
            const int sharp = di + dj - ta - tb;
            const int lam = dmax;
            const Closing c = closing_value(dmax, dmin, tri[p], 1, sharp, lam);
            if (sharp_out) sharp_out[p] = sharp;
            if (lam_out) lam_out[p] = lam;
            if (c64_out) c64_out[p] = c.c64;
            c32_out[p] = c.c32;
        }
    }
}

// Asymmetric 0/1 adjacency (directed simple graph): the loop of bfc_cuda.py:31-44 evaluated sparsely by the
// definition (dcr_directed.cuh).  One warp per entry i -> j of the successor CSR.
__global__ void __launch_bounds__(256) cuda_flavour_directed_kernel(GraphView out, GraphView in, int n,
                                                                    int32_t* __restrict__ tri_out,
                                                                    int32_t* __restrict__ sharp_out,
                                                                    int32_t* __restrict__ lam_out,
                                                                    double* __restrict__ c64_out,
                                                                    float* __restrict__ c32_out, int64_t lo, int64_t hi) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t p = lo + (((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5); p < hi; p += warps) {
        const int i = row_of_entry(out.start, n, p);
        const int j = out.col[p];
        int sharp, lam, a2;
        const Closing c = directed_entry_curvature(out, in, i, j, lane, &sharp, &lam, &a2);
        if (lane == 0) {
            if (tri_out) tri_out[p] = a2;
            if (sharp_out) sharp_out[p] = sharp;
            if (lam_out) lam_out[p] = lam;
            if (c64_out) c64_out[p] = c.c64;
            c32_out[p] = c.c32;
        }
    }
}

}  // namespace dcr

using namespace dcr;

static inline int grid_for_warps(int64_t items) {
    // 256-thread CTAs = 8 warps; cap at 8 resident CTAs per SM, grid a multiple of the SM count
    const int sms = sm_count();
    int64_t ctas = (items + 7) / 8;
    const int64_t cap = (int64_t)sms * 8;
    if (ctas > cap) ctas = cap;
    if (ctas < 1) ctas = 1;
    return (int)ctas;
}

extern "C" int dcr_bfc_support(const int32_t* rowptr, const int32_t* colidx, int n, int32_t* tri, int64_t entry_lo,
                               int64_t entry_hi, void* stream) {
    if (entry_hi <= entry_lo) return 0;
    support_kernel<<<grid_for_warps(entry_hi - entry_lo), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n, tri,
                                                                                         entry_lo, entry_hi);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_bfc_cuda_flavour(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* tri,
                                    int32_t* sharp, int32_t* lam, double* c64, float* c32, int64_t entry_lo,
                                    int64_t entry_hi, void* stream) {
    if (entry_hi <= entry_lo) return 0;
    if (!c32) { set_error("dcr_bfc_cuda_flavour: c32 must not be NULL"); return 1; }
    cuda_flavour_kernel<<<grid_for_warps(entry_hi - entry_lo), 256, 0, (cudaStream_t)stream>>>(
        rowptr, colidx, n, tri, sharp, lam, c64, c32, entry_lo, entry_hi);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_bfc_cuda_flavour_directed(const int32_t* out_rowptr, const int32_t* out_colidx,
                                             const int32_t* in_rowptr, const int32_t* in_colidx, int n, int32_t* tri,
                                             int32_t* sharp, int32_t* lam, double* c64, float* c32, int64_t entry_lo,
                                             int64_t entry_hi, void* stream) {
    if (entry_hi <= entry_lo) return 0;
    if (!c32) { set_error("dcr_bfc_cuda_flavour_directed: c32 must not be NULL"); return 1; }
    GraphView out{out_rowptr, nullptr, out_colidx}, in{in_rowptr, nullptr, in_colidx};
    cuda_flavour_directed_kernel<<<grid_for_warps(entry_hi - entry_lo), 256, 0, (cudaStream_t)stream>>>(
        out, in, n, tri, sharp, lam, c64, c32, entry_lo, entry_hi);
    DCR_LAUNCH_CHECK();
    return 0;
}
