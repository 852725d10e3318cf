// dcr_score.cuh — candidate scoring shared by dcr_post_delta (static CSR) and the SDRF loop (dynamic arena).
//
// Takes over _balanced_forman_post_delta (curvature/bfc_cuda.py:68-141): D[I,J] = cuda-flavour curvature of (x,y)
// on A + e_i e_j^T for i = i_nb[I], j = j_nb[J].  The reference launches one thread per cell with an N-long z
// loop over dense A and A@A.  Only z in N(x) ∪ N(y) ∪ {i,j} contribute, and the cell-independent
// part of each term is shared by all cells, so one CTA does:
//   1. base1[q] = A2[x,z] - A[x,z] for z = q-th neighbour of y,  base2[p] = A2[z,y] - A[z,y] for z = p-th
//      neighbour of x  (supports are looked up, not recomputed, when the pair is an edge; A2[v,v] = d_v);
//   2. sharp0 / lam0 = the count of positive base terms and their maximum;
//   3. per cell: the patches of bfc_cuda.py:113-124 touch one z (cells with x != i and y != j) -> O(1) per cell,
//      one thread per cell; cells with x == i or y == j patch every z -> one warp per cell.
// All integer arithmetic is exact; the closing formula is dcr::closing_value (fp64, two fp32 roundings).
// Requires symmetric 0/1 adjacency without self-loops (is_undirected=True, the only mode the reference's callers
// use: rewiring/rewire.py:10, ph/eval_rewiring_ph.py:34).
#pragma once

#include "dcr_common.cuh"

namespace dcr {

constexpr float MASKED_D = -1000.0f;  // bfc_cuda.py:78

struct ScoreScratch {
    int32_t* base1;  // [deg(y)]
    int32_t* base2;  // [deg(x)]
    int32_t* posI;   // [n_i]  position of i_nb[I] in row x (absolute slot) or -1
    int32_t* posJ;   // [n_j]  position of j_nb[J] in row y or -1
};

struct ScoreShared {   // block-shared state of one scoring call
    int axy, dx, dy, sx, sy, a2xy, sharp0, lam0;
    ClosingFirst first;      // closing formula up to the first fp32 store, for the unchanged degrees (d_x, d_y)
};

__device__ __forceinline__ int edge_slot(const GraphView& g, int a, int b) {
    return find_sorted(g.col, g.begin(a), g.degree(a), b);
}

// A2[a,b] for arbitrary nodes: d_a on the diagonal, the stored support for an edge, an intersection otherwise.
// Warp-cooperative (all 32 lanes call it with the same arguments).
__device__ __forceinline__ int warp_a2(const GraphView& g, const int32_t* supp, int a, int b, int lane) {
    if (a == b) return g.degree(a);
    const int s = edge_slot(g, a, b);
    if (s >= 0) return supp[s];
    return warp_intersect_count(g, a, b, lane);
}

// Steps 1-2.  Called by every thread of the CTA; `sh` lives in shared memory.  n_i/n_j lists are accessed through
// the functors nbI(I) / nbJ(J) so the SDRF loop can serve "insertion-order row + [self]" without materialising it.
template <class NbI, class NbJ>
__device__ void score_prepare(const GraphView& g, const int32_t* supp, int x, int y, NbI nbI, int n_i, NbJ nbJ,
                              int n_j, const ScoreScratch& sc, ScoreShared* sh) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    if (tid == 0) {
        sh->dx = g.degree(x);
        sh->dy = g.degree(y);
        sh->sx = g.begin(x);
        sh->sy = g.begin(y);
        const int s = (x == y) ? -1 : edge_slot(g, x, y);
        sh->axy = s >= 0;
        sh->a2xy = s >= 0 ? supp[s] : 0;   // only used multiplied by A[x,y]
        sh->sharp0 = 0;
        sh->lam0 = 0;
        sh->first.c32 = 0.0f;
        sh->first.dmax = 1.0;
        if (sh->dx > 0 && sh->dy > 0)
            sh->first = closing_first(max(sh->dx, sh->dy), min(sh->dx, sh->dy), sh->a2xy, sh->axy);
    }
    __syncthreads();
    const int dx = sh->dx, dy = sh->dy, sx = sh->sx, sy = sh->sy;
    for (int I = tid; I < n_i; I += nthreads) sc.posI[I] = find_sorted(g.col, sx, dx, nbI(I));
    for (int J = tid; J < n_j; J += nthreads) sc.posJ[J] = find_sorted(g.col, sy, dy, nbJ(J));
    if (!sh->axy) { __syncthreads(); return; }   // every TMP is multiplied by A[x,y] = 0 (bfc_cuda.py:127,133)
    int cnt = 0, mx = 0;
    for (int q = warp; q < dy; q += nwarps) {    // T1 terms: z in N(y)
        const int z = g.col[sy + q];
        const int b = warp_a2(g, supp, x, z, lane) - (z != x && edge_slot(g, x, z) >= 0 ? 1 : 0);
        if (lane == 0) { sc.base1[q] = b; cnt += b > 0; mx = max(mx, b); }
    }
    for (int p = warp; p < dx; p += nwarps) {    // T2 terms: z in N(x)
        const int z = g.col[sx + p];
        const int b = warp_a2(g, supp, z, y, lane) - (z != y && edge_slot(g, z, y) >= 0 ? 1 : 0);
        if (lane == 0) { sc.base2[p] = b; cnt += b > 0; mx = max(mx, b); }
    }
    if (lane == 0) {
        if (cnt) atomicAdd(&sh->sharp0, cnt);
        if (mx) atomicMax(&sh->lam0, mx);
    }
    __syncthreads();
}

// One cell with x != i and y != j (thread-level).  Returns D[I,J].
__device__ __forceinline__ float score_cell_simple(const GraphView& g, const ScoreScratch& sc, const ScoreShared* sh,
                                                   int x, int y, int i, int j, int I, int J) {
    if (i == j || edge_slot(g, i, j) >= 0) return MASKED_D;           // bfc_cuda.py:77-79
    int din = sh->dx, dout = sh->dy;
    if (j == x) din += 1; else if (i == y) dout += 1;                 // :82-85
    if (din == 0 || dout == 0) return 0.0f;                           // :87-89
    const int dmax = max(din, dout), dmin = min(din, dout);           // :91-96
    const bool bumped = (j == x) || (i == y);                         // never true for an unmasked cell of SDRF
    if (!sh->axy) return bumped ? closing_value(dmax, dmin, 0, 0, 0, 0).c32 : sh->first.c32;
    const int pI = sc.posI[I], pJ = sc.posJ[J];
    int sharp = sh->sharp0, lam = sh->lam0;
    if (pJ >= 0 && pI >= 0) {          // z == j: A2_x_z += A[x,i]   (:123-124)
        const int b = sc.base1[pJ - sh->sy];
        sharp += (b <= 0);             // b + 1 > 0 whenever b >= 0
        lam = max(lam, b + 1);
    }
    if (pI >= 0 && pJ >= 0) {          // z == i: A2_z_y += A[j,y]   (:117-118)
        const int b = sc.base2[pI - sh->sx];
        sharp += (b <= 0);
        lam = max(lam, b + 1);
    }
    // :139-141 (no triangle patch: x != i, y != j); the first store is shared by all such cells
    if (bumped) return closing_value(dmax, dmin, sh->a2xy, 1, sharp, lam).c32;
    return closing_finish(sh->first, sharp, lam);
}

// One cell with x == i or y == j (warp-level; all lanes pass identical arguments).  Returns D[I,J].
__device__ __forceinline__ float score_cell_warp(const GraphView& g, const int32_t* supp, const ScoreScratch& sc,
                                                 const ScoreShared* sh, int x, int y, int i, int j, int lane) {
    if (i == j || edge_slot(g, i, j) >= 0) return MASKED_D;
    int din = sh->dx, dout = sh->dy;
    if (j == x) din += 1; else if (i == y) dout += 1;
    if (din == 0 || dout == 0) return 0.0f;
    const int dmax = max(din, dout), dmin = min(din, dout);
    if (!sh->axy) return closing_value(dmax, dmin, 0, 0, 0, 0).c32;
    const int dx = sh->dx, dy = sh->dy, sx = sh->sx, sy = sh->sy;
    const int axi = (x == i) ? 0 : (edge_slot(g, x, i) >= 0);   // A[x,i]
    const int ajy = (j == y) ? 0 : (edge_slot(g, j, y) >= 0);   // A[j,y]
    int a2 = sh->a2xy;                                          // :99-103
    if (x == i && ajy) a2 += 1; else if (y == j && axi) a2 += 1;
    int cnt = 0, mx = 0;
    for (int q = lane; q < dy; q += 32) {                       // z in N(y): A_z_y = 1
        const int z = g.col[sy + q];
        int v = sc.base1[q];
        if (x == i) v += (z != j && edge_slot(g, j, z) >= 0);   // :119-120  A2_x_z += A[j,z]
        if (z == j) v += axi;                                   // :123-124
        if (x == i && z == j) v -= 1;                           // :115-116  A_x_z += 1
        cnt += v > 0;
        mx = max(mx, v);
    }
    for (int p = lane; p < dx; p += 32) {                       // z in N(x): A_x_z = 1
        const int z = g.col[sx + p];
        int v = sc.base2[p];
        if (z == i) v += ajy;                                   // :117-118
        if (y == j) v += (z != i && edge_slot(g, z, i) >= 0);   // :121-122  A2_z_y += A[z,i]
        if (z == i && y == j) v -= 1;                           // :113-114  A_z_y += 1
        cnt += v > 0;
        mx = max(mx, v);
    }
    cnt = warp_sum(cnt);
    mx = warp_max(mx);
    // the one z outside N(x) ∪ N(y) that the patched A makes non-zero
    if (y == j && x != i) {          // z == i: A_z_y = 0 + 1 ; T1 = A2[x,i] (+0) - A[x,i]
        const int v = warp_a2(g, supp, x, i, lane) - axi;
        cnt += v > 0;
        mx = max(mx, v);
    }
    if (x == i && y != j) {          // z == j: A_x_z = 0 + 1 ; T2 = A2[j,y] - A[j,y]
        const int v = warp_a2(g, supp, j, y, lane) - ajy;
        cnt += v > 0;
        mx = max(mx, v);
    }
    return closing_value(dmax, dmin, a2, 1, cnt, mx).c32;
}

// Step 3 for the whole matrix.  `out(I, J, d)` receives every cell value.
template <class NbI, class NbJ, class Out>
__device__ void score_cells(const GraphView& g, const int32_t* supp, int x, int y, NbI nbI, int n_i, NbJ nbJ,
                            int n_j, const ScoreScratch& sc, const ScoreShared* sh, Out out) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarps = nthreads >> 5;
    const long long cells = (long long)n_i * n_j;
    for (long long c = tid; c < cells; c += nthreads) {
        const int I = (int)(c / n_j), J = (int)(c - (long long)I * n_j);
        const int i = nbI(I), j = nbJ(J);
        if (i == x || j == y) continue;
        out(I, J, score_cell_simple(g, sc, sh, x, y, i, j, I, J));
    }
    // rows with i == x and columns with j == y: one warp per cell
    for (int I = 0; I < n_i; ++I) {
        if (nbI(I) != x) continue;
        for (int J = warp; J < n_j; J += nwarps) {
            const float d = score_cell_warp(g, supp, sc, sh, x, y, x, nbJ(J), lane);
            if (lane == 0) out(I, J, d);
        }
    }
    for (int J = 0; J < n_j; ++J) {
        if (nbJ(J) != y) continue;
        for (int I = warp; I < n_i; I += nwarps) {
            if (nbI(I) == x) continue;   // done above
            const float d = score_cell_warp(g, supp, sc, sh, x, y, nbI(I), y, lane);
            if (lane == 0) out(I, J, d);
        }
    }
}

}  // namespace dcr
