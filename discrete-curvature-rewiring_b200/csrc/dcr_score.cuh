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
    // aggregates of the base terms, so that a cell which patches only a few terms never re-walks all of them:
    int cnt1, cnt2;          // #{base1 > 0}, #{base2 > 0}
    int m1, i1, m1x;         // max base1, an index attaining it, max over the OTHER indices (-1 = none)
    int m2, i2, m2x;         // same for base2
};

// (max, an index attaining it, max over the other indices, #positive) of arr[0..n) — warp-cooperative, every lane
// returns the same values
__device__ __forceinline__ void warp_top2(const int32_t* arr, int n, int lane, int& m, int& idx, int& mx, int& pos) {
    m = -1; idx = -1; mx = -1; pos = 0;
    for (int q = lane; q < n; q += 32) {
        const int v = arr[q];
        pos += v > 0;
        if (v > m) { mx = m; m = v; idx = q; } else mx = max(mx, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const int om = __shfl_xor_sync(FULL, m, o), oi = __shfl_xor_sync(FULL, idx, o), ox = __shfl_xor_sync(FULL, mx, o);
        if (om > m) { mx = max(m, ox); m = om; idx = oi; } else mx = max(mx, om);
    }
    pos = warp_sum(pos);
    m = __shfl_sync(FULL, m, 0); idx = __shfl_sync(FULL, idx, 0); mx = __shfl_sync(FULL, mx, 0);
}

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

// A2[a,b] - A[a,b] with one adjacency search (the base term of bfc_cuda.py:127 / :133)
__device__ __forceinline__ int warp_a2_minus_a(const GraphView& g, const int32_t* supp, int a, int b, int lane) {
    if (a == b) return g.degree(a);
    const int s = edge_slot(g, a, b);
    if (s >= 0) return supp[s] - 1;
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
        sh->cnt1 = sh->cnt2 = 0;
        sh->m1 = sh->m1x = sh->m2 = sh->m2x = -1;
        sh->i1 = sh->i2 = -1;
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
        const int b = warp_a2_minus_a(g, supp, x, z, lane);
        if (lane == 0) { sc.base1[q] = b; cnt += b > 0; mx = max(mx, b); }
    }
    for (int p = warp; p < dx; p += nwarps) {    // T2 terms: z in N(x)
        const int z = g.col[sx + p];
        const int b = warp_a2_minus_a(g, supp, z, y, lane);
        if (lane == 0) { sc.base2[p] = b; cnt += b > 0; mx = max(mx, b); }
    }
    (void)cnt; (void)mx;
    __syncthreads();
    if (warp < 2) {          // warp 0: aggregates of base1, warp 1 (or warp 0 again): aggregates of base2
        for (int which = warp; which < 2; which += nwarps >= 2 ? 2 : 1) {
            int m, idx, mxo, pos;
            warp_top2(which == 0 ? sc.base1 : sc.base2, which == 0 ? dy : dx, lane, m, idx, mxo, pos);
            if (lane == 0) {
                if (which == 0) { sh->cnt1 = pos; sh->m1 = m; sh->i1 = idx; sh->m1x = mxo; }
                else            { sh->cnt2 = pos; sh->m2 = m; sh->i2 = idx; sh->m2x = mxo; }
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        sh->sharp0 = sh->cnt1 + sh->cnt2;
        sh->lam0 = max(0, max(sh->m1, sh->m2));
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
// With x == i the patches of bfc_cuda.py:113-124 read, for every z in N(y):  v(z) = base1(z) + A[j,z] - [z == j];
// the T2 terms over N(x) are unchanged, and one extra T2 term appears at z = j.  So instead of re-walking all of
// N(y) with a membership test per element, only the COMMON neighbours of j and y matter:
//     #positive = cnt1 + #{w in N(j)∩N(y) : base1(w) = 0} - [j in N(y) and base1(j) = 1]
//     max       = max( max base1 over z != j,  max over common w of base1(w) + 1,  base1(j) - 1 )
// (base terms are >= 0 whenever A[x,y] = 1, and base1(j) >= 1 for j in N(y)).  y == j is the mirror image with
// N(x), base2 and the extra T1 term at z = i.  All integer and exact.
__device__ __forceinline__ float score_cell_warp(const GraphView& g, const int32_t* supp, const ScoreScratch& sc,
                                                 const ScoreShared* sh, int x, int y, int i, int j, int lane) {
    if (i == j || edge_slot(g, i, j) >= 0) return MASKED_D;
    int din = sh->dx, dout = sh->dy;
    if (j == x) din += 1; else if (i == y) dout += 1;
    if (din == 0 || dout == 0) return 0.0f;
    const int dmax = max(din, dout), dmin = min(din, dout);
    if (!sh->axy) return closing_value(dmax, dmin, 0, 0, 0, 0).c32;
    // orient: `r` = the endpoint of (x,y) whose row is patched (y when i == x, x when j == y), `o` = the new
    // neighbour on the other side (j resp. i); base/aggregates of that side in (base, cntP, mP, iP, mPx)
    const bool rowA = (x == i);                    // class A: i == x, patch the N(y) terms with N(j)
    const int o = rowA ? j : i;                    // (r = y when i == x, r = x when j == y)
    const int sr = rowA ? sh->sy : sh->sx, dr = rowA ? sh->dy : sh->dx;
    const int32_t* base = rowA ? sc.base1 : sc.base2;
    const int cntP = rowA ? sh->cnt1 : sh->cnt2, cntQ = rowA ? sh->cnt2 : sh->cnt1;
    const int mP = rowA ? sh->m1 : sh->m2, iP = rowA ? sh->i1 : sh->i2, mPx = rowA ? sh->m1x : sh->m2x;
    const int mQ = rowA ? sh->m2 : sh->m1;
    const int so = g.begin(o), d_o = g.degree(o);
    // common neighbours of o and r, located in r's row
    int czero = 0, mmark = -1, ncommon = 0;
    if (d_o <= dr) {
        for (int t = lane; t < d_o; t += 32) {
            const int q = find_sorted(g.col, sr, dr, g.col[so + t]);
            if (q >= 0) { const int b = base[q - sr]; ++ncommon; czero += (b == 0); mmark = max(mmark, b + 1); }
        }
    } else {
        for (int q = lane; q < dr; q += 32) {
            if (find_sorted(g.col, so, d_o, g.col[sr + q]) >= 0) {
                const int b = base[q]; ++ncommon; czero += (b == 0); mmark = max(mmark, b + 1);
            }
        }
    }
    czero = warp_sum(czero);
    ncommon = warp_sum(ncommon);
    mmark = warp_max(mmark);
    const int po = find_sorted(g.col, sr, dr, o);            // is o itself a neighbour of r (A[j,y] resp. A[x,i])?
    const int aor = po >= 0;
    int cnt = cntP + czero, mx = max(mmark, mQ);
    if (po >= 0) {
        const int b = base[po - sr];                         // >= 1: r's partner endpoint is a common neighbour
        cnt -= (b == 1);
        mx = max(mx, b - 1);
        mx = max(mx, (po - sr) == iP ? mPx : mP);
    } else {
        mx = max(mx, mP);
    }
    cnt += cntQ;
    // the one z outside N(x) ∪ N(y) that the patched A makes non-zero: z = o, value A2[o,r] - A[o,r]
    const int extra = (po >= 0 ? supp[po] : ncommon) - aor;
    cnt += extra > 0;
    mx = max(mx, extra);
    const int a2 = sh->a2xy + aor;                           // bfc_cuda.py:99-103
    return closing_value(dmax, dmin, a2, 1, cnt, max(mx, 0)).c32;
}

// Step 3 for the whole matrix.  `out(I, J, d)` receives every cell value.
// `only_I` / `only_J` >= 0: the caller knows that x (resp. y) occurs exactly once in the list, at that position
// (the SDRF loop appends it last, sdrf_cuda_bfc.py:45-46) — saves every thread a walk over the lists.
template <class NbI, class NbJ, class Out>
__device__ void score_cells(const GraphView& g, const int32_t* supp, int x, int y, NbI nbI, int n_i, NbJ nbJ,
                            int n_j, const ScoreScratch& sc, const ScoreShared* sh, Out out, int only_I = -1,
                            int only_J = -1, int part = 0, int parts = 1) {
    // (part, parts): this CTA's share when several CTAs score one matrix (dcr_post_delta on large candidate sets)
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int lane = tid & 31, nwarps = (nthreads >> 5) * parts, warp = part * (nthreads >> 5) + (tid >> 5);
    const long long cells = (long long)n_i * n_j;
    for (long long c = (long long)part * nthreads + tid; c < cells; c += (long long)parts * nthreads) {
        const int I = (int)(c / n_j), J = (int)(c - (long long)I * n_j);
        const int i = nbI(I), j = nbJ(J);
        if (i == x || j == y) continue;
        out(I, J, score_cell_simple(g, sc, sh, x, y, i, j, I, J));
    }
    // rows with i == x and columns with j == y: one warp per cell
    for (int I = (only_I >= 0 ? only_I : 0); I < (only_I >= 0 ? only_I + 1 : n_i); ++I) {
        if (nbI(I) != x) continue;
        for (int J = warp; J < n_j; J += nwarps) {
            const float d = score_cell_warp(g, supp, sc, sh, x, y, x, nbJ(J), lane);
            if (lane == 0) out(I, J, d);
        }
    }
    for (int J = (only_J >= 0 ? only_J : 0); J < (only_J >= 0 ? only_J + 1 : n_j); ++J) {
        if (nbJ(J) != y) continue;
        for (int I = warp; I < n_i; I += nwarps) {
            if (nbI(I) == x) continue;   // done above
            const float d = score_cell_warp(g, supp, sc, sh, x, y, nbI(I), y, lane);
            if (lane == 0) out(I, J, d);
        }
    }
}

}  // namespace dcr
