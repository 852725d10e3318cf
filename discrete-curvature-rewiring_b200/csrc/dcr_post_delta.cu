// dcr_post_delta.cu — the dense-signature candidate scoring entry points over a static CSR.
// Takes over balanced_forman_post_delta (curvature/bfc_cuda.py:144-159); the arithmetic lives in dcr_score.cuh (symmetric
// A) and dcr_directed.cuh (asymmetric A).  Two launches: ONE CTA computes what all cells share (base terms over N(x), N(y)
// and their aggregates, positions of the list entries) into the caller's workspace, then as many CTAs as the candidate
// matrix warrants (one per 4096 cells, up to two per SM) score their share of the cells.  The caller owns every byte:
// the workspace of dcr_post_delta_workspace_bytes(n, n_i, n_j) bytes replaces the stream-ordered allocation of round 1.
#include <algorithm>

#include "dcr_directed.cuh"

namespace dcr {

constexpr int PD_THREADS = 512;

template <class Shared>
struct PdWorkspace {
    Shared* shared;
    ScoreScratch sc;
};
template <class Shared>
static PdWorkspace<Shared> pd_layout(void* workspace, int n, int n_i, int n_j) {
    PdWorkspace<Shared> w;
    char* p = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    w.shared = (Shared*)p; p += 256;
    int32_t* ints = (int32_t*)p;
    w.sc.base1 = ints;
    w.sc.base2 = ints + n;
    w.sc.posI = ints + 2 * (size_t)n;
    w.sc.posJ = w.sc.posI + n_i;
    return w;
}

__global__ void __launch_bounds__(PD_THREADS) post_delta_prepare_kernel(GraphView g, const int32_t* __restrict__ supp, int x,
                                                                        int y, const int32_t* __restrict__ i_nb, int n_i,
                                                                        const int32_t* __restrict__ j_nb, int n_j,
                                                                        ScoreScratch sc, ScoreShared* out) {
    __shared__ ScoreShared sh;
    score_prepare(g, supp, x, y, [=](int I) { return i_nb[I]; }, n_i, [=](int J) { return j_nb[J]; }, n_j, sc, &sh);
    __syncthreads();
    if (threadIdx.x == 0) *out = sh;
}
__global__ void __launch_bounds__(PD_THREADS) post_delta_cells_kernel(GraphView g, const int32_t* __restrict__ supp, int x,
                                                                      int y, const int32_t* __restrict__ i_nb, int n_i,
                                                                      const int32_t* __restrict__ j_nb, int n_j,
                                                                      ScoreScratch sc, const ScoreShared* in,
                                                                      float* __restrict__ D) {
    __shared__ ScoreShared sh;
    if (threadIdx.x == 0) sh = *in;
    __syncthreads();
    score_cells(g, supp, x, y, [=](int I) { return i_nb[I]; }, n_i, [=](int J) { return j_nb[J]; }, n_j, sc, &sh,
                [=](int I, int J, float d) { D[(size_t)I * n_j + J] = d; }, -1, -1, (int)blockIdx.x, (int)gridDim.x);
}

// asymmetric A: successors `out`, predecessors `in` (dcr_directed.cuh)
__global__ void __launch_bounds__(PD_THREADS) post_delta_directed_prepare_kernel(GraphView out, GraphView in, int x, int y,
                                                                                 const int32_t* __restrict__ i_nb, int n_i,
                                                                                 const int32_t* __restrict__ j_nb, int n_j,
                                                                                 ScoreScratch sc, DirScoreShared* res) {
    __shared__ DirScoreShared sh;
    __shared__ int red[2];
    directed_score_prepare(out, in, x, y, [=](int I) { return i_nb[I]; }, n_i, [=](int J) { return j_nb[J]; }, n_j, sc, &sh,
                           red);
    __syncthreads();
    if (threadIdx.x == 0) *res = sh;
}
__global__ void __launch_bounds__(PD_THREADS) post_delta_directed_cells_kernel(GraphView out, GraphView in, int x, int y,
                                                                               const int32_t* __restrict__ i_nb, int n_i,
                                                                               const int32_t* __restrict__ j_nb, int n_j,
                                                                               ScoreScratch sc, const DirScoreShared* src,
                                                                               float* __restrict__ D) {
    __shared__ DirScoreShared sh;
    if (threadIdx.x == 0) sh = *src;
    __syncthreads();
    directed_score_cells(out, in, x, y, [=](int I) { return i_nb[I]; }, n_i, [=](int J) { return j_nb[J]; }, n_j, sc, &sh,
                         [=](int I, int J, float d) { D[(size_t)I * n_j + J] = d; }, (int)blockIdx.x, (int)gridDim.x);
}

}  // namespace dcr

using namespace dcr;

extern "C" int64_t dcr_post_delta_workspace_bytes(int n, int n_i, int n_j) {
    return 512 + ((int64_t)2 * n + n_i + n_j) * (int64_t)sizeof(int32_t);
}

static int pd_grid(int n_i, int n_j) {
    const long long cells = (long long)n_i * n_j;
    return (int)std::max<long long>(1, std::min<long long>((cells + 4095) / 4096, (long long)sm_count() * 2));
}

static int pd_check(const char* who, int n, int x, int y, int n_i, int n_j, void* workspace, int64_t workspace_bytes) {
    if (x < 0 || y < 0 || x >= n || y >= n) { set_error("%s: (x,y) out of range", who); return 1; }
    if (!workspace || workspace_bytes < dcr_post_delta_workspace_bytes(n, n_i, n_j)) {
        set_error("%s: workspace of dcr_post_delta_workspace_bytes(n, n_i, n_j) bytes required", who);
        return 1;
    }
    return 0;
}

extern "C" int dcr_post_delta(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* tri, int x, int y,
                              const int32_t* i_nb, int n_i, const int32_t* j_nb, int n_j, float* D, void* workspace,
                              int64_t workspace_bytes, void* stream) {
    if (n_i <= 0 || n_j <= 0) return 0;
    if (pd_check("dcr_post_delta", n, x, y, n_i, n_j, workspace, workspace_bytes)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const PdWorkspace<ScoreShared> w = pd_layout<ScoreShared>(workspace, n, n_i, n_j);
    GraphView g{rowptr, nullptr, colidx};
    post_delta_prepare_kernel<<<1, PD_THREADS, 0, st>>>(g, tri, x, y, i_nb, n_i, j_nb, n_j, w.sc, w.shared);
    DCR_LAUNCH_CHECK();
    post_delta_cells_kernel<<<pd_grid(n_i, n_j), PD_THREADS, 0, st>>>(g, tri, x, y, i_nb, n_i, j_nb, n_j, w.sc, w.shared, D);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_post_delta_directed(const int32_t* out_rowptr, const int32_t* out_colidx, const int32_t* in_rowptr,
                                       const int32_t* in_colidx, int n, int x, int y, const int32_t* i_nb, int n_i,
                                       const int32_t* j_nb, int n_j, float* D, void* workspace, int64_t workspace_bytes,
                                       void* stream) {
    if (n_i <= 0 || n_j <= 0) return 0;
    if (pd_check("dcr_post_delta_directed", n, x, y, n_i, n_j, workspace, workspace_bytes)) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    const PdWorkspace<DirScoreShared> w = pd_layout<DirScoreShared>(workspace, n, n_i, n_j);
    GraphView out{out_rowptr, nullptr, out_colidx}, in{in_rowptr, nullptr, in_colidx};
    post_delta_directed_prepare_kernel<<<1, PD_THREADS, 0, st>>>(out, in, x, y, i_nb, n_i, j_nb, n_j, w.sc, w.shared);
    DCR_LAUNCH_CHECK();
    post_delta_directed_cells_kernel<<<pd_grid(n_i, n_j), PD_THREADS, 0, st>>>(out, in, x, y, i_nb, n_i, j_nb, n_j, w.sc,
                                                                               w.shared, D);
    DCR_LAUNCH_CHECK();
    return 0;
}
