// dcr_directed.cuh — cuda-flavour BFC and candidate scoring for an ASYMMETRIC 0/1 adjacency (is_undirected=False).
//
// Takes over _balanced_forman_curvature (curvature/bfc_cuda.py:11-48) and _balanced_forman_post_delta (:68-141) for
// a directed simple graph, where the closed form of the symmetric case (lambda = d_max, sharp from the supports,
// dcr_bfc_cuda.cu) no longer holds: A2[i,k] = |N_out(i) ∩ N_in(k)| differs from A2[k,i], the degrees are the
// IN-degree of i and the OUT-degree of j (:20-25, :54-55), and the loop terms of :33-44
//     T1(k) = A[k,j] (A2[i,k] - A[i,k])   -> k in N_in(j):   |N_out(i) ∩ N_in(k)| - [i -> k]
//     T2(k) = A[i,k] (A2[k,j] - A[k,j])   -> k in N_out(i):  |N_out(k) ∩ N_in(j)| - [k -> j]
// can be -1, 0 or positive.  The N-long loop of the reference runs over dense rows; here only the k that make the
// leading factor non-zero are visited, each term is one sorted-list intersection, all integer and exact, and the
// closing formula is dcr::closing_value (the compiled kernel's fp64 dataflow with two fp32 roundings).
// The graph is given as two views with sorted rows: `out` (successors) and `in` (predecessors).
#pragma once

#include "dcr_score.cuh"

namespace dcr {

// |A ∩ B| of two sorted ranges, by one thread: elements of the shorter range are searched in the longer one
__device__ __forceinline__ int thread_intersect(const int32_t* __restrict__ ca, int sa, int da,
                                                const int32_t* __restrict__ cb, int sb, int db) {
    if (da > db) {
        const int32_t* tc = ca; ca = cb; cb = tc;
        int t = sa; sa = sb; sb = t;
        t = da; da = db; db = t;
    }
    int c = 0;
    for (int t = 0; t < da; ++t) c += find_sorted(cb, sb, db, ca[sa + t]) >= 0;
    return c;
}

// A2[a,b] = #{m : a -> m -> b}
__device__ __forceinline__ int directed_a2(const GraphView& out, const GraphView& in, int a, int b) {
    return thread_intersect(out.col, out.begin(a), out.degree(a), in.col, in.begin(b), in.degree(b));
}
__device__ __forceinline__ int directed_has(const GraphView& out, int a, int b) {   // A[a,b]
    return find_sorted(out.col, out.begin(a), out.degree(a), b) >= 0;
}

// C[i,j] of an existing entry i -> j (bfc_cuda.py:15-48).  Warp-cooperative: lanes take the loop's k.
__device__ __forceinline__ Closing directed_entry_curvature(const GraphView& out, const GraphView& in, int i, int j,
                                                            int lane, int* sharp_out = nullptr,
                                                            int* lam_out = nullptr, int* a2_out = nullptr) {
    const int din = in.degree(i), dout = out.degree(j);        // d_in[i], d_out[j]  (:20-25)
    Closing zero;
    zero.c64 = 0.0;
    zero.c32 = 0.0f;
    const int so = out.begin(i), no = out.degree(i);
    const int si = in.begin(j), ni = in.degree(j);
    int a2 = 0;                                                 // A2[i,j]
    for (int t = lane; t < no; t += 32) a2 += find_sorted(in.col, si, ni, out.col[so + t]) >= 0;
    a2 = warp_sum(a2);
    if (sharp_out) *sharp_out = 0;
    if (lam_out) *lam_out = 0;
    if (a2_out) *a2_out = a2;
    if (din == 0 || dout == 0) return zero;                     // :27-29
    int sharp = 0, lam = 0;
    for (int t = lane; t < ni; t += 32) {                       // T1: k -> j
        const int k = in.col[si + t];
        const int v = directed_a2(out, in, i, k) - (find_sorted(out.col, so, no, k) >= 0);
        if (v > 0) { ++sharp; lam = max(lam, v); }
    }
    for (int t = lane; t < no; t += 32) {                       // T2: i -> k
        const int k = out.col[so + t];
        const int v = directed_a2(out, in, k, j) - (find_sorted(in.col, si, ni, k) >= 0);
        if (v > 0) { ++sharp; lam = max(lam, v); }
    }
    sharp = warp_sum(sharp);
    lam = warp_max(lam);
    if (sharp_out) *sharp_out = sharp;
    if (lam_out) *lam_out = lam;
    return closing_value(max(din, dout), min(din, dout), a2, 1, sharp, lam);
}

// ------------------------------------------------------------------------------------------------------------
// candidate scoring: D[I,J] = C[x,y] on A + e_i e_j^T  (bfc_cuda.py:68-141)
// ------------------------------------------------------------------------------------------------------------
struct DirScoreShared {
    int axy, din, dout, sox, nox, siy, niy, a2xy, sharp0, lam0;
};

// base1[q] = A2[x,z] - A[x,z] for z = q-th predecessor of y (sorted), base2[p] = A2[z,y] - A[z,y] for z = p-th
// successor of x; posI[I] / posJ[J] = index of i_nb[I] in the successors of x / of j_nb[J] in the predecessors of y.
template <class NbI, class NbJ>
__device__ void directed_score_prepare(const GraphView& out, const GraphView& in, int x, int y, NbI nbI, int n_i,
                                       NbJ nbJ, int n_j, const ScoreScratch& sc, DirScoreShared* sh, int* red) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    if (tid == 0) {
        sh->din = in.degree(x);                                 // A[:,x].sum()  (:147)
        sh->dout = out.degree(y);                               // A[y].sum()    (:148)
        sh->sox = out.begin(x); sh->nox = out.degree(x);
        sh->siy = in.begin(y);  sh->niy = in.degree(y);
        sh->axy = (x != y) && directed_has(out, x, y);
        sh->a2xy = sh->axy ? directed_a2(out, in, x, y) : 0;    // only used multiplied by A[x,y]
        sh->sharp0 = 0;
        sh->lam0 = 0;
        red[0] = 0;
        red[1] = 0;
    }
    __syncthreads();
    const int sox = sh->sox, nox = sh->nox, siy = sh->siy, niy = sh->niy;
    for (int I = tid; I < n_i; I += nthreads) {
        const int p = find_sorted(out.col, sox, nox, nbI(I));
        sc.posI[I] = p >= 0 ? p - sox : -1;
    }
    for (int J = tid; J < n_j; J += nthreads) {
        const int q = find_sorted(in.col, siy, niy, nbJ(J));
        sc.posJ[J] = q >= 0 ? q - siy : -1;
    }
    if (!sh->axy) { __syncthreads(); return; }                  // every TMP carries the factor A[x,y] (:127,:133)
    int cnt = 0, mx = 0;
    for (int q = tid; q < niy; q += nthreads) {
        const int z = in.col[siy + q];
        const int b = directed_a2(out, in, x, z) - (find_sorted(out.col, sox, nox, z) >= 0);
        sc.base1[q] = b;
        cnt += b > 0;
        mx = max(mx, b);
    }
    for (int p = tid; p < nox; p += nthreads) {
        const int z = out.col[sox + p];
        const int b = directed_a2(out, in, z, y) - (find_sorted(in.col, siy, niy, z) >= 0);
        sc.base2[p] = b;
        cnt += b > 0;
        mx = max(mx, b);
    }
    cnt = warp_sum(cnt);
    mx = warp_max(mx);
    if ((tid & 31) == 0) { atomicAdd(&red[0], cnt); atomicMax(&red[1], mx); }
    __syncthreads();
    if (tid == 0) { sh->sharp0 = red[0]; sh->lam0 = red[1]; }
    __syncthreads();
}

__device__ __forceinline__ bool directed_cell_degrees(const DirScoreShared* sh, int x, int y, int i, int j, int& dmax,
                                                      int& dmin) {
    int din = sh->din, dout = sh->dout;
    if (j == x) din += 1; else if (i == y) dout += 1;           // :82-85
    dmax = max(din, dout);
    dmin = min(din, dout);
    return din != 0 && dout != 0;                               // :87-89
}

// cell with x != i and y != j: the patches of :113-124 touch z = j (T1) and z = i (T2) only — O(1), one thread
__device__ __forceinline__ float directed_cell_simple(const GraphView& out, const ScoreScratch& sc,
                                                      const DirScoreShared* sh, int x, int y, int i, int j, int I,
                                                      int J) {
    if (i == j || directed_has(out, i, j)) return MASKED_D;     // :77-79
    int dmax, dmin;
    if (!directed_cell_degrees(sh, x, y, i, j, dmax, dmin)) return 0.0f;
    if (!sh->axy) return closing_value(dmax, dmin, 0, 0, 0, 0).c32;
    const int pI = sc.posI[I], pJ = sc.posJ[J];
    int sharp = sh->sharp0, lam = sh->lam0;
    if (pI >= 0 && pJ >= 0) {
        const int b1 = sc.base1[pJ];                            // z == j: A2_x_z += A[x,i]   (:123-124)
        sharp += (b1 == 0);
        lam = max(lam, b1 + 1);
        const int b2 = sc.base2[pI];                            // z == i: A2_z_y += A[j,y]   (:117-118)
        sharp += (b2 == 0);
        lam = max(lam, b2 + 1);
    }
    return closing_value(dmax, dmin, sh->a2xy, 1, sharp, lam).c32;
}

// any cell, by the definition (:99-141): one warp, lanes take the z that make a leading factor non-zero
__device__ __forceinline__ float directed_cell_warp(const GraphView& out, const GraphView& in, const ScoreScratch& sc,
                                                    const DirScoreShared* sh, int x, int y, int i, int j, int lane) {
    if (i == j || directed_has(out, i, j)) return MASKED_D;
    int dmax, dmin;
    if (!directed_cell_degrees(sh, x, y, i, j, dmax, dmin)) return 0.0f;
    if (!sh->axy) return closing_value(dmax, dmin, 0, 0, 0, 0).c32;
    const int sox = sh->sox, nox = sh->nox, siy = sh->siy, niy = sh->niy;
    const bool xi = (x == i), yj = (y == j);
    const int a_xi = find_sorted(out.col, sox, nox, i) >= 0;    // A[x,i]
    const int a_jy = find_sorted(in.col, siy, niy, j) >= 0;     // A[j,y]
    int a2 = sh->a2xy;                                          // :99-103
    if (xi && a_jy) a2 += 1; else if (yj && a_xi) a2 += 1;
    int sharp = 0, lam = 0;
    // T1 = A_z_y (A2_x_z - A_x_z): z -> y, plus z = i when the new entry is i -> y
    for (int q = lane; q < niy; q += 32) {
        const int z = in.col[siy + q];
        int v = sc.base1[q];
        if (xi) v += directed_has(out, j, z) - (z == j);        // + A[j,z] (:119-120), A_x_z += 1 (:115-116)
        if (z == j) v += a_xi;                                  // :123-124
        if (v > 0) { ++sharp; lam = max(lam, v); }
    }
    // T2 = A_x_z (A2_z_y - A_z_y): x -> z, plus z = j when the new entry is x -> j
    for (int p = lane; p < nox; p += 32) {
        const int z = out.col[sox + p];
        int v = sc.base2[p];
        if (yj) v += directed_has(out, z, i) - (z == i);        // + A[z,i] (:121-122), A_z_y += 1 (:113-114)
        if (z == i) v += a_jy;                                  // :117-118
        if (v > 0) { ++sharp; lam = max(lam, v); }
    }
    if (lane == 0) {
        if (yj && find_sorted(in.col, siy, niy, i) < 0) {       // z = i now points at y: A2[x,i] - A[x,i] (+ A[j,i] if x == i)
            int v = directed_a2(out, in, x, i) - a_xi;
            if (xi) v += directed_has(out, j, i);
            if (v > 0) { ++sharp; lam = max(lam, v); }
        }
        if (xi && find_sorted(out.col, sox, nox, j) < 0) {      // x now points at z = j: A2[j,y] - A[j,y] (+ A[j,i] if y == j)
            int v = directed_a2(out, in, j, y) - a_jy;
            if (yj) v += directed_has(out, j, i);
            if (v > 0) { ++sharp; lam = max(lam, v); }
        }
    }
    sharp = warp_sum(sharp);
    lam = warp_max(lam);
    return closing_value(dmax, dmin, a2, 1, sharp, lam).c32;
}

template <class NbI, class NbJ, class Out>
__device__ void directed_score_cells(const GraphView& out, const GraphView& in, int x, int y, NbI nbI, int n_i, NbJ nbJ,
                                     int n_j, const ScoreScratch& sc, const DirScoreShared* sh, Out put, int part = 0,
                                     int parts = 1) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, nwarps = (nthreads >> 5) * parts, warp = part * (nthreads >> 5) + (tid >> 5);
    const long long cells = (long long)n_i * n_j;
    for (long long c = (long long)part * nthreads + tid; c < cells; c += (long long)parts * nthreads) {
        const int I = (int)(c / n_j), J = (int)(c - (long long)I * n_j);
        const int i = nbI(I), j = nbJ(J);
        if (i == x || j == y) continue;
        put(I, J, directed_cell_simple(out, sc, sh, x, y, i, j, I, J));
    }
    for (long long c = warp; c < cells; c += nwarps) {          // cells with i == x or j == y
        const int I = (int)(c / n_j), J = (int)(c - (long long)I * n_j);
        const int i = nbI(I), j = nbJ(J);
        if (i != x && j != y) continue;
        const float d = directed_cell_warp(out, in, sc, sh, x, y, i, j, lane);
        if (lane == 0) put(I, J, d);
    }
}

}  // namespace dcr
