// dcr_dense_small.cu — cuda-flavour BFC straight from a SMALL dense adjacency (n <= 1024), for the reference's dense
// signature balanced_forman_curvature(A, C) (curvature/bfc_cuda.py:51-65) on the WebKB-sized graphs.
//
// On a 183- or 251-node graph the CSR route is launch-bound (count, scan, fill, supports, closing, scatter + a host
// round trip for nnz: ~0.17 ms against 0.12 ms for the reference's one numba kernel).  Here the whole matrix is
// bit-packed once (n/32 words per row: the graph lives in a few KB of L1/L2), and ONE kernel — a CTA per row —
// computes, for every entry of the row that sits on an edge, the support |N(i) ∩ N(j)| as popc(row_i & row_j), the two
// "support == 1" counts over the common neighbours (App. A.2) the same way, and the closing formula of the compiled
// reference kernel; the other entries of the row are written as +0.0, so C needs no memset.  Two launches; when the
// pack kernel finds A outside the covered domain (flags != 0) the second kernel writes nothing.
#include "dcr_common.cuh"

namespace dcr {

constexpr int SMALL_MAX_N = 1024, SMALL_MAX_WORDS = SMALL_MAX_N / 32;

// one warp per row: bit-pack, degree, validation (0/1 values, zero diagonal, symmetry)
__global__ void dense_small_pack_kernel(const float* __restrict__ A, int n, int words, uint32_t* __restrict__ bits,
                                        int32_t* __restrict__ deg, int32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int row = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const float* r = A + (size_t)row * n;
    int cnt = 0, bad = 0;
    for (int w = 0; w < words; ++w) {
        const int c = 32 * w + lane;
        const float v = (c < n) ? r[c] : 0.0f;
        if (v != 0.0f) {
            if (v != 1.0f) bad |= 1;
            if (c == row) bad |= 2;
            if (A[(size_t)c * n + row] != v) bad |= 4;
        }
        const unsigned m = __ballot_sync(FULL, v != 0.0f);
        cnt += __popc(m);
        if (lane == 0) bits[(size_t)row * words + w] = m;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) bad |= __shfl_xor_sync(FULL, bad, o);
    if (lane == 0) {
        deg[row] = cnt;
        if (bad) atomicOr(flags, bad);
    }
}

// |N(u) ∩ N(v)| from the packed rows; every lane gets the result (words <= 32: one word per lane)
__device__ __forceinline__ int packed_support(const uint32_t* ru, const uint32_t* __restrict__ rv, int words, int lane) {
    const int c = lane < words ? __popc(ru[lane] & rv[lane]) : 0;
    return __reduce_add_sync(FULL, c);
}

// CTA per row i (4 warps): warps take the entries j of the row in turn
__global__ void __launch_bounds__(128) dense_small_bfc_kernel(int n, int words, const uint32_t* __restrict__ bits,
                                                              const int32_t* __restrict__ deg,
                                                              const int32_t* __restrict__ flags, float* __restrict__ C) {
    if (*flags) return;   // outside the covered domain (pack kernel's verdict): leave the caller's C untouched
    __shared__ uint32_t s_row[SMALL_MAX_WORDS];
    __shared__ uint32_t s_common[4][SMALL_MAX_WORDS];
    const int i = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < words) s_row[threadIdx.x] = bits[(size_t)i * words + threadIdx.x];
    __syncthreads();
    const int di = deg[i];
    float* out = C + (size_t)i * n;
    for (int j = warp; j < n; j += 4) {
        if (!((s_row[j >> 5] >> (j & 31)) & 1u)) {        // off-edge (and diagonal): curvature/bfc_cuda.py:16-18
            if (lane == 0) out[j] = 0.0f;
            continue;
        }
        const uint32_t* rj = bits + (size_t)j * words;
        const uint32_t cw = lane < words ? (s_row[lane] & rj[lane]) : 0u;   // common neighbours, one word per lane
        const int a2 = __reduce_add_sync(FULL, __popc(cw));
        int t1 = 0, t2 = 0;                                // #{k common : c(i,k) == 1}, #{k common : c(k,j) == 1}
        if (a2 > 0) {
            s_common[warp][lane] = cw;
            __syncwarp();
            for (int w = 0; w < words; ++w) {
                uint32_t m = s_common[warp][w];
                while (m) {
                    const int k = 32 * w + __ffs(m) - 1;
                    m &= m - 1;
                    const uint32_t* rk = bits + (size_t)k * words;
                    t1 += packed_support(s_row, rk, words, lane) == 1;
                    t2 += packed_support(rj, rk, words, lane) == 1;
                }
            }
            __syncwarp();
        }
        if (lane == 0) {
            const int dj = deg[j];
            const int dmax = max(di, dj), dmin = min(di, dj);
            // symmetric 0/1 A without self-loops: sharp = d_i + d_j - t1 - t2, lambda = d_max
            out[j] = closing_value(dmax, dmin, a2, 1, (long long)di + dj - t1 - t2, dmax).c32;
        }
    }
}

}  // namespace dcr

using namespace dcr;

extern "C" int64_t dcr_bfc_cuda_dense_small_workspace_bytes(int n) {
    const int words = (n + 31) / 32;
    return (int64_t)n * words * (int64_t)sizeof(uint32_t) + (int64_t)n * (int64_t)sizeof(int32_t);
}

extern "C" int dcr_bfc_cuda_dense_small(const float* A, int n, float* C, int32_t* flags, void* workspace,
                                        int64_t workspace_bytes, void* stream) {
    if (n <= 0) return 0;
    if (n > SMALL_MAX_N) { set_error("dcr_bfc_cuda_dense_small: n = %d above %d", n, SMALL_MAX_N); return 1; }
    if (workspace_bytes < dcr_bfc_cuda_dense_small_workspace_bytes(n)) {
        set_error("dcr_bfc_cuda_dense_small: workspace too small");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int words = (n + 31) / 32;
    uint32_t* bits = (uint32_t*)workspace;
    int32_t* deg = (int32_t*)(bits + (size_t)n * words);
    dense_small_pack_kernel<<<(unsigned)(((size_t)n * 32 + 127) / 128), 128, 0, st>>>(A, n, words, bits, deg, flags);
    DCR_LAUNCH_CHECK();
    dense_small_bfc_kernel<<<n, 128, 0, st>>>(n, words, bits, deg, flags, C);
    DCR_LAUNCH_CHECK();
    return 0;
}
