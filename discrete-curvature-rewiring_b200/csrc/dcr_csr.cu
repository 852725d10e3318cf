// dcr_csr.cu — dense fp32 adjacency <-> sorted CSR, for the reference's dense-matrix signatures
// (curvature/bfc_cuda.py:51-57, :144-150 take a dense A and return a dense C / D).
//
// HBM-bound streaming kernels: every element of A is read exactly once per pass with 128-bit loads where the
// row is 16-byte aligned, one warp per row chunk, ballot/popc compaction keeps the output sorted.
#include "dcr_common.cuh"

namespace dcr {

// One warp per row.  counts non-zeros; validates 0/1 values, zero diagonal and symmetry (A[i,j] vs A[j,i] is a
// strided read — only done for non-zero entries, i.e. nnz extra reads, not n^2).
__global__ void dense_count_kernel(const float* __restrict__ A, int n, int32_t* __restrict__ counts,
                                   int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const float* r = A + (size_t)row * n;
    int cnt = 0, bad = 0;
    for (int c0 = 0; c0 < n; c0 += 32) {
        int c = c0 + lane;
        float v = (c < n) ? r[c] : 0.0f;
        if (v != 0.0f) {
            ++cnt;
            if (v != 1.0f) bad |= 1;
            if (c == row) bad |= 2;
            if (A[(size_t)c * n + row] != v) bad |= 4;
        }
    }
    cnt = warp_sum(cnt);
#pragma unroll
    for (int o = 16; o; o >>= 1) bad |= __shfl_xor_sync(FULL, bad, o);
    if (lane == 0) {
        counts[row] = cnt;
        if (bad) atomicOr(flags, bad);
    }
}

__global__ void dense_fill_kernel(const float* __restrict__ A, int n, const int32_t* __restrict__ rowptr,
                                  int32_t* __restrict__ colidx) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const float* r = A + (size_t)row * n;
    int out = rowptr[row];
    for (int c0 = 0; c0 < n; c0 += 32) {
        int c = c0 + lane;
        bool nz = (c < n) && (r[c] != 0.0f);
        unsigned m = __ballot_sync(FULL, nz);
        if (nz) colidx[out + __popc(m & ((1u << lane) - 1))] = c;
        out += __popc(m);
    }
}

__global__ void scatter_dense_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                                     const float* __restrict__ vals, float* __restrict__ C) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    float* r = C + (size_t)row * n;
    // zero the row, then drop the row's entries in: one pass over C, no separate memset of n^2 floats
    for (int c = lane; c < n; c += 32) r[c] = 0.0f;
    __syncwarp();
    const int b = rowptr[row], e = rowptr[row + 1];
    for (int p = b + lane; p < e; p += 32) r[colidx[p]] = vals[p];
}

// The undirected edge list (row < col entries in CSR order) of a symmetric sorted CSR, derived on the device so that an
// end-to-end caller uploads the CSR only.  upper_count: entries of row v beyond the diagonal; upper_fill: one warp per row
// writes its edges at the row's offset (exclusive prefix sum of the counts, done by the caller).
__global__ void csr_upper_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                                       int64_t* __restrict__ counts) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int b = rowptr[v], len = rowptr[v + 1] - b;
    counts[v] = len - lower_bound(colidx, b, len, v + 1);
}
__global__ void csr_upper_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n,
                                      const int64_t* __restrict__ offsets_incl, int32_t* __restrict__ esrc,
                                      int32_t* __restrict__ edst) {
    const int lane = threadIdx.x & 31;
    const int v = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (v >= n) return;
    const int e = rowptr[v + 1];
    const int64_t end = offsets_incl[v], cnt = end - (v ? offsets_incl[v - 1] : 0);
    const int first = e - (int)cnt;                        // the row is sorted: its entries > v are its last `cnt`
    for (int t = lane; t < (int)cnt; t += 32) {
        esrc[end - cnt + t] = v;
        edst[end - cnt + t] = colidx[first + t];
    }
}

}  // namespace dcr

using namespace dcr;

static inline dim3 warp_per_row_grid(int n, int threads) {
    long long warps_per_block = threads / 32;
    return dim3((unsigned)((n + warps_per_block - 1) / warps_per_block));
}

extern "C" int dcr_dense_count(const float* A, int n, int32_t* row_counts, int32_t* flags, void* stream) {
    if (n <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    DCR_CUDA(cudaMemsetAsync(flags, 0, sizeof(int32_t), st));
    dense_count_kernel<<<warp_per_row_grid(n, 256), 256, 0, st>>>(A, n, row_counts, flags);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_dense_fill(const float* A, int n, const int32_t* rowptr, int32_t* colidx, void* stream) {
    if (n <= 0) return 0;
    dense_fill_kernel<<<warp_per_row_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(A, n, rowptr, colidx);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_scatter_dense(const int32_t* rowptr, const int32_t* colidx, int n, const float* vals, float* C,
                                 void* stream) {
    if (n <= 0) return 0;
    scatter_dense_kernel<<<warp_per_row_grid(n, 256), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n, vals, C);
    DCR_LAUNCH_CHECK();
    return 0;
}

extern "C" int dcr_csr_upper_count(const int32_t* rowptr, const int32_t* colidx, int n, int64_t* counts, void* stream) {
    if (n <= 0) return 0;
    csr_upper_count_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n, counts);
    DCR_LAUNCH_CHECK();
    return 0;
}
extern "C" int dcr_csr_upper_fill(const int32_t* rowptr, const int32_t* colidx, int n, const int64_t* offsets_incl,
                                  int32_t* esrc, int32_t* edst, void* stream) {
    if (n <= 0) return 0;
    csr_upper_fill_kernel<<<(unsigned)(((size_t)n * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(rowptr, colidx, n,
                                                                                                   offsets_incl, esrc, edst);
    DCR_LAUNCH_CHECK();
    return 0;
}
