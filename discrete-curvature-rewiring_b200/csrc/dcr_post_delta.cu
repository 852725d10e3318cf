// dcr_post_delta.cu — the dense-signature candidate scoring entry point over a static CSR.
// Takes over balanced_forman_post_delta (curvature/bfc_cuda.py:144-159); the arithmetic lives in dcr_score.cuh.
#include "dcr_directed.cuh"

namespace dcr {

__global__ void __launch_bounds__(512) post_delta_kernel(GraphView g, const int32_t* __restrict__ supp, int x, int y,
                                                         const int32_t* __restrict__ i_nb, int n_i,
                                                         const int32_t* __restrict__ j_nb, int n_j,
                                                         ScoreScratch sc, float* __restrict__ D) {
    __shared__ ScoreShared sh;
    auto nbI = [=](int I) { return i_nb[I]; };
    auto nbJ = [=](int J) { return j_nb[J]; };
    score_prepare(g, supp, x, y, nbI, n_i, nbJ, n_j, sc, &sh);
    score_cells(g, supp, x, y, nbI, n_i, nbJ, n_j, sc, &sh,
                [=](int I, int J, float d) { D[(size_t)I * n_j + J] = d; });
}

// asymmetric A: successors `out`, predecessors `in` (dcr_directed.cuh)
__global__ void __launch_bounds__(512) post_delta_directed_kernel(GraphView out, GraphView in, int x, int y,
                                                                  const int32_t* __restrict__ i_nb, int n_i,
                                                                  const int32_t* __restrict__ j_nb, int n_j,
                                                                  ScoreScratch sc, float* __restrict__ D) {
    __shared__ DirScoreShared sh;
    __shared__ int red[2];
    auto nbI = [=](int I) { return i_nb[I]; };
    auto nbJ = [=](int J) { return j_nb[J]; };
    directed_score_prepare(out, in, x, y, nbI, n_i, nbJ, n_j, sc, &sh, red);
    directed_score_cells(out, in, x, y, nbI, n_i, nbJ, n_j, sc, &sh,
                         [=](int I, int J, float d) { D[(size_t)I * n_j + J] = d; });
}

}  // namespace dcr

using namespace dcr;

extern "C" int dcr_post_delta(const int32_t* rowptr, const int32_t* colidx, int n, const int32_t* tri, int x, int y,
                              const int32_t* i_nb, int n_i, const int32_t* j_nb, int n_j, float* D, void* stream) {
    if (n_i <= 0 || n_j <= 0) return 0;
    if (x < 0 || y < 0 || x >= n || y >= n) { set_error("dcr_post_delta: (x,y) out of range"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    // scratch: base1[deg y] + base2[deg x] <= 2n ints, posI[n_i], posJ[n_j]
    int32_t* buf = nullptr;
    const size_t ints = (size_t)2 * n + n_i + n_j;
    DCR_CUDA(cudaMallocAsync((void**)&buf, ints * sizeof(int32_t), st));
    ScoreScratch sc;
    sc.base1 = buf;
    sc.base2 = buf + n;
    sc.posI = buf + 2 * (size_t)n;
    sc.posJ = sc.posI + n_i;
    GraphView g{rowptr, nullptr, colidx};
    post_delta_kernel<<<1, 512, 0, st>>>(g, tri, x, y, i_nb, n_i, j_nb, n_j, sc, D);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(buf, st);
    if (e != cudaSuccess) return cuda_fail(e, "post_delta_kernel", __FILE__, __LINE__);
    return 0;
}

extern "C" int dcr_post_delta_directed(const int32_t* out_rowptr, const int32_t* out_colidx, const int32_t* in_rowptr,
                                       const int32_t* in_colidx, int n, int x, int y, const int32_t* i_nb, int n_i,
                                       const int32_t* j_nb, int n_j, float* D, void* stream) {
    if (n_i <= 0 || n_j <= 0) return 0;
    if (x < 0 || y < 0 || x >= n || y >= n) { set_error("dcr_post_delta_directed: (x,y) out of range"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    int32_t* buf = nullptr;
    const size_t ints = (size_t)2 * n + n_i + n_j;
    DCR_CUDA(cudaMallocAsync((void**)&buf, ints * sizeof(int32_t), st));
    ScoreScratch sc;
    sc.base1 = buf;
    sc.base2 = buf + n;
    sc.posI = buf + 2 * (size_t)n;
    sc.posJ = sc.posI + n_i;
    GraphView out{out_rowptr, nullptr, out_colidx}, in{in_rowptr, nullptr, in_colidx};
    post_delta_directed_kernel<<<1, 512, 0, st>>>(out, in, x, y, i_nb, n_i, j_nb, n_j, sc, D);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(buf, st);
    if (e != cudaSuccess) return cuda_fail(e, "post_delta_directed_kernel", __FILE__, __LINE__);
    return 0;
}
