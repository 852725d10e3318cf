// dcr_tc.cu — dense-regime supports on the 5th-generation tensor cores: A2 = A·A with tcgen05 kind::i8.
//
// Takes over the `A2 = torch.matmul(A, A)` of the reference (curvature/bfc_cuda.py:53, :146) for graphs small and
// dense enough that the product really is a dense contraction (N <= 32768: A as int8 is <= 1 GiB).  For a 0/1
// symmetric adjacency A2[i,j] = |N(i) ∩ N(j)| is exact in int32.  The N x N product is NEVER materialised: the
// epilogue of every 128x128 tile keeps only the entries that sit on an edge and writes them straight into the
// CSR-ordered support array that dcr_bfc_cuda_flavour / dcr_post_delta / the SDRF state consume (same output as
// dcr_bfc_support, which is the sparse route for the same numbers).
//
// Structure (one 128x128 output tile per CTA, K swept in 128-byte blocks):
//   warp 0   TMA producer   cp.async.bulk.tensor.2d of the A-rows tile (128 x 128 B) and the "B" tile — by symmetry
//                           also 128 rows of A, K-major — into a 3-stage 128B-swizzled shared-memory ring,
//                           mbarrier complete_tx
//   warp 1   MMA issuer     one elected thread: 4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N128 K32) per stage,
//                           accumulators in TMEM (128 lanes x 128 columns of int32); tcgen05.commit frees the stage
//   warps 2-5 epilogue      tcgen05.ld 32x32b.x32 (each warp its 32-lane quarter), then per row: walk the 128-byte
//                           row segment of A and store acc[c] for every non-zero A[i, j] at its CSR position
// A·A is symmetric: only the upper-triangular tiles are computed (half the MMAs); a coalesced pass over the CSR then
// copies every value to the reverse entry below the block diagonal.
// Roofline: tensor-bound, 2·N_pad³ int8 ops.  Evidence: UTCIMMA / UTMALDG / LDTM in the SASS (cuobjdump), and
// sm__pipe_tensor_cycles_active under ncu.
#include <cuda.h>

#include "dcr_common.cuh"

namespace dcr {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 128;      // tile (int8 elements; BK bytes = one 128B swizzle row)
constexpr int TC_STAGES = 3;                              // 3 x 32 KB: two CTAs per SM, one's epilogue under the other's MMAs
constexpr int TC_UMMA_K = 32;                             // int8: 32 elements = 32 bytes per MMA
constexpr int TC_THREADS = 192;                           // 6 warps: TMA, MMA, 4 x epilogue
constexpr int TC_TMEM_COLS = 128;
constexpr int TC_STAGE_BYTES = TC_BM * TC_BK;             // 16 KB per operand per stage
constexpr int TC_SMEM_BYTES = 2 * TC_STAGES * TC_STAGE_BYTES + 1024;   // + alignment slack

// ---- thin PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int x, int y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    const uint32_t zero = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(zero) : "memory");
}

// Shared-memory matrix descriptor, K-major, 128B swizzle (cute::UMMA::SmemDescriptor): start address >> 4 in bits
// [0,14), leading byte offset (unused for swizzled K-major: 1) in [16,30), stride byte offset = 8 rows x 128 B = 1024
// >> 4 in [32,46), version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 @[4,6), a/b format INT8 = 1 @[7,10)/[10,13),
// a/b K-major = 0 @15/@16, N>>3 @[17,23), M>>4 @[24,29).
__host__ __device__ constexpr uint32_t umma_idesc_i8(int m, int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ---- set-up kernels ----------------------------------------------------------------------------------------
// int8 image of the adjacency (zero-padded to n_pad) from the CSR, and per (row, 128-column block) the number of
// the row's entries before that block (so an epilogue thread knows where its block's entries start in the CSR).
__global__ void tc_fill_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx, int n, int n_pad,
                               int8_t* __restrict__ A8, int32_t* __restrict__ blkpre) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const int b = rowptr[row], e = rowptr[row + 1];
    for (int p = b + lane; p < e; p += 32) {
        const int c = colidx[p];
        A8[(size_t)row * n_pad + c] = 1;
    }
    const int nblk = n_pad / TC_BN;
    for (int t = lane; t < nblk; t += 32) blkpre[(size_t)row * nblk + t] = lower_bound(colidx, b, e - b, t * TC_BN);
}

// entries below the block diagonal of a symmetric product: A2[i,j] = A2[j,i].  One thread per directed entry (a warp
// per row would leave a hub row's 3420 binary searches to 32 lanes).
__global__ void __launch_bounds__(256) tc_mirror_kernel(const int32_t* __restrict__ rowptr,
                                                        const int32_t* __restrict__ colidx, int n, int64_t nnz,
                                                        int32_t* __restrict__ tri) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int a = 0, b = n;                       // row of entry p: rowptr[a] <= p < rowptr[b]
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if ((int64_t)rowptr[m] <= p) a = m; else b = m;
    }
    const int c = colidx[p];
    if (c / TC_BN >= a / TC_BM) return;     // on or above the block diagonal: written by the product kernel
    const int cb = rowptr[c];
    tri[p] = tri[find_sorted(colidx, cb, rowptr[c + 1] - cb, a)];
}

// ---- the GEMM + fused edge-extraction kernel ---------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS, 2)
tc_support_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const int8_t* __restrict__ A8,
                  const int32_t* __restrict__ rowptr, const int32_t* __restrict__ blkpre, int n, int n_pad,
                  int32_t* __restrict__ tri, int symmetric) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[TC_STAGES], empty_bar[TC_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_slot;

    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);   // 128B swizzle: 1024-byte aligned
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + TC_STAGES * TC_STAGE_BYTES;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // symmetric != 0: the product is SYMMETRIC (A·A) — only the tiles with nt >= mt are computed (1-D grid over the
    // upper triangle, row by row); tc_mirror_kernel copies the values to the entries below the block diagonal
    int mt = blockIdx.y, nt = blockIdx.x;
    if (symmetric) {
        const int T = n_pad / TC_BM;
        int t = blockIdx.x;
        mt = 0;
        while (t >= T - mt) { t -= T - mt; ++mt; }
        nt = mt + t;
    }
    const int num_kb = n_pad / TC_BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
        for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {   // TMEM allocation is a warp-wide instruction; the same warp frees it
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t phase = (kb / TC_STAGES) & 1;
                mbar_wait(&empty_bar[s], phase ^ 1);
                mbar_expect_tx(&full_bar[s], 2 * TC_STAGE_BYTES);
                tma_load_2d(smem_a + s * TC_STAGE_BYTES, &tmap_a, &full_bar[s], kb * TC_BK, mt * TC_BM);
                tma_load_2d(smem_b + s * TC_STAGE_BYTES, &tmap_b, &full_bar[s], kb * TC_BK, nt * TC_BN);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_i8(TC_BM, TC_BN);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % TC_STAGES;
                const uint32_t phase = (kb / TC_STAGES) & 1;
                mbar_wait(&full_bar[s], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem_a + s * TC_STAGE_BYTES);
                const uint32_t b_addr = smem_u32(smem_b + s * TC_STAGE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
                    umma_i8(tmem_base, umma_desc_k_sw128(a_addr + k * TC_UMMA_K), umma_desc_k_sw128(b_addr + k * TC_UMMA_K),
                            idesc, (kb | k) != 0);
                }
                umma_commit(&empty_bar[s]);          // frees the stage once these MMAs have read it
            }
            umma_commit(&tmem_full_bar);             // accumulators complete
        }
    } else {
        // epilogue: warp w owns TMEM lanes 32*(w % 4) .. +31 = rows of the tile
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int i = mt * TC_BM + r;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int nblk = n_pad / TC_BN;
        int pos = (i < n) ? rowptr[i] + blkpre[(size_t)i * nblk + nt] : 0;
#pragma unroll 1
        for (int c0 = 0; c0 < TC_BN; c0 += 32) {
            uint32_t acc[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(acc[0]), "=r"(acc[1]), "=r"(acc[2]), "=r"(acc[3]), "=r"(acc[4]), "=r"(acc[5]), "=r"(acc[6]),
                  "=r"(acc[7]), "=r"(acc[8]), "=r"(acc[9]), "=r"(acc[10]), "=r"(acc[11]), "=r"(acc[12]), "=r"(acc[13]),
                  "=r"(acc[14]), "=r"(acc[15]), "=r"(acc[16]), "=r"(acc[17]), "=r"(acc[18]), "=r"(acc[19]),
                  "=r"(acc[20]), "=r"(acc[21]), "=r"(acc[22]), "=r"(acc[23]), "=r"(acc[24]), "=r"(acc[25]),
                  "=r"(acc[26]), "=r"(acc[27]), "=r"(acc[28]), "=r"(acc[29]), "=r"(acc[30]), "=r"(acc[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (i < n) {
                const uint4* seg = (const uint4*)(A8 + (size_t)i * n_pad + (size_t)nt * TC_BN + c0);   // 32 bytes
                const uint4 lo = seg[0], hi = seg[1];
                const uint32_t w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    if ((w[c >> 2] >> ((c & 3) * 8)) & 0xffu) tri[pos++] = (int32_t)acc[c];
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

// Measurement aid: the int8 tensor rate this GPU sustains with the SAME instruction the product kernel issues
// (tcgen05.mma.cta_group::1.kind::i8, M128 N128 K32, operands resident in shared memory, accumulator in TMEM), no loads
// and no epilogue — the denominator bench.py holds the product kernel against (MEASURED_PEAKS.json has no int8 figure).
__global__ void __launch_bounds__(64, 2) tc_int8_peak_kernel(int iters) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t done_bar;
    __shared__ uint32_t tmem_base_slot;
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int p = threadIdx.x; p < 2 * TC_STAGE_BYTES / 4; p += blockDim.x) ((uint32_t*)smem)[p] = 0x01010101u;
    if (warp == 0 && lane == 0) {
        mbar_init(&done_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tmem_base_slot)), "r"((uint32_t)TC_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy stores -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    if (warp == 0 && lane == 0) {
        constexpr uint32_t idesc = umma_idesc_i8(TC_BM, TC_BN);
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + TC_STAGE_BYTES);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
                umma_i8(tmem_base, umma_desc_k_sw128(a_addr + k * TC_UMMA_K), umma_desc_k_sw128(b_addr + k * TC_UMMA_K), idesc,
                        (it | k) != 0);
        }
        umma_commit(&done_bar);
        mbar_wait(&done_bar, 0);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TC_TMEM_COLS)
                     : "memory");
    }
}

}  // namespace dcr

using namespace dcr;

// TOP/s (int8, dense) of the resident-tile loop; synchronises.  *out_tops is DEVICE memory (fp64).
extern "C" int dcr_tc_int8_peak(double* out_tops, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * TC_STAGE_BYTES + 1024, iters = 2048, ctas = sm_count() * 2;
    DCR_CUDA(cudaFuncSetAttribute(tc_int8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    DCR_CUDA(cudaEventCreate(&e0));
    DCR_CUDA(cudaEventCreate(&e1));
    tc_int8_peak_kernel<<<ctas, 64, smem, st>>>(64);                    // warm-up
    DCR_CUDA(cudaEventRecord(e0, st));
    tc_int8_peak_kernel<<<ctas, 64, smem, st>>>(iters);
    DCR_CUDA(cudaEventRecord(e1, st));
    DCR_CUDA(cudaEventSynchronize(e1));
    float ms = 0.0f;
    DCR_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double ops = 2.0 * TC_BM * TC_BN * TC_BK * (double)iters * ctas;
    const double tops = ops / (ms * 1e-3) / 1e12;
    DCR_CUDA(cudaMemcpyAsync(out_tops, &tops, sizeof(double), cudaMemcpyHostToDevice, st));
    DCR_CUDA(cudaStreamSynchronize(st));
    return 0;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)p;
    }
    return fn;
}

static inline int tc_pad(int n) { return (n + TC_BM - 1) / TC_BM * TC_BM; }

static int make_tmap(CUtensorMap* tmap, const int8_t* base, int np) {
    PFN_encodeTiled encode = get_encode_tiled();
    if (!encode) { set_error("tensor path: cuTensorMapEncodeTiled unavailable"); return 1; }
    const cuuint64_t gdim[2] = {(cuuint64_t)np, (cuuint64_t)np};           // inner (columns, bytes), outer (rows)
    const cuuint64_t gstride[1] = {(cuuint64_t)np};                        // bytes between rows
    const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_BM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)base, gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("tensor path: cuTensorMapEncodeTiled failed (%d)", (int)r); return 1; }
    return 0;
}

static int launch_edge_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const int8_t* A8, const int32_t* rowptr,
                            const int32_t* blkpre, int n, int np, int32_t* out, bool symmetric, cudaStream_t st) {
    static bool attr_done_dev[MAX_DEVICES] = {false};
    bool& attr_done = attr_done_dev[current_device()];
    if (!attr_done) {
        DCR_CUDA(cudaFuncSetAttribute(tc_support_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES));
        attr_done = true;
    }
    const unsigned T = (unsigned)(np / TC_BM);
    const dim3 grid = symmetric ? dim3(T * (T + 1) / 2, 1) : dim3(T, T);
    tc_support_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, st>>>(ta, tb, A8, rowptr, blkpre, n, np, out, symmetric ? 1 : 0);
    DCR_LAUNCH_CHECK();
    return 0;
}

static int launch_mirror(const int32_t* rowptr, const int32_t* colidx, int n, int64_t nnz, int32_t* tri, cudaStream_t st) {
    tc_mirror_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(rowptr, colidx, n, nnz, tri);
    DCR_LAUNCH_CHECK();
    return 0;
}

struct TcWorkspace {
    int8_t* A8;
    int8_t* Q8;
    int32_t* blkpre;
    int32_t* t1;
    size_t total;
};
static TcWorkspace tc_layout(void* workspace, int n, int64_t nnz, bool with_q) {
    TcWorkspace w;
    const size_t np = (size_t)tc_pad(n);
    char* p = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    char* p0 = p;
    w.A8 = (int8_t*)p; p += np * np;
    w.Q8 = (int8_t*)p; if (with_q) p += np * np;
    w.blkpre = (int32_t*)p; p += ((size_t)n * (np / TC_BN) * sizeof(int32_t) + 255) / 256 * 256;
    w.t1 = (int32_t*)p; if (with_q) p += ((size_t)nnz * sizeof(int32_t) + 255) / 256 * 256;
    w.total = (size_t)(p - p0) + 256;
    return w;
}

extern "C" int64_t dcr_bfc_support_tc_workspace_bytes(int n, int64_t nnz) {
    return (int64_t)tc_layout(nullptr, n, nnz, false).total;
}

extern "C" int dcr_bfc_support_tc(const int32_t* rowptr, const int32_t* colidx, int n, int64_t nnz, int32_t* tri,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    if (n <= 0 || nnz <= 0) return 0;
    if (n > 32768) { set_error("dcr_bfc_support_tc: dense tensor-core path is for n <= 32768 (got %d)", n); return 1; }
    if (workspace_bytes < dcr_bfc_support_tc_workspace_bytes(n, nnz)) { set_error("dcr_bfc_support_tc: workspace too small"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int np = tc_pad(n);
    const TcWorkspace w = tc_layout(workspace, n, nnz, false);
    DCR_CUDA(cudaMemsetAsync(w.A8, 0, (size_t)np * np, st));
    tc_fill_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(rowptr, colidx, n, np, w.A8, w.blkpre);
    DCR_LAUNCH_CHECK();
    CUtensorMap tmap;
    if (make_tmap(&tmap, w.A8, np)) return 1;
    if (launch_edge_gemm(tmap, tmap, w.A8, rowptr, w.blkpre, n, np, tri, true, st)) return 1;   // symmetric: upper triangle
    return launch_mirror(rowptr, colidx, n, nnz, tri, st);
}

// ------------------------------------------------------------------------------------------------------------
// The whole cuda flavour in the dense regime: two tensor-core products and an elementwise closing pass.
//   A2 = A·A          -> tri[p] = A2[i,j] on the edges                               (bfc_cuda.py:53)
//   T1 = Q·A, Q = A ∧ [A2 == 1]  -> t1[p] = #{k in N(i)∩N(j) : support(i,k) == 1}   (the zero terms of :34-44)
//   sharp = d_i + d_j - t1[i,j] - t1[j,i],  lambda = d_max                            
// ------------------------------------------------------------------------------------------------------------
namespace dcr {
__global__ void tc_fill_q_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                                 const int32_t* __restrict__ tri, int n, int n_pad, int8_t* __restrict__ Q8) {
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= n) return;
    const int b = rowptr[row], e = rowptr[row + 1];
    for (int p = b + lane; p < e; p += 32)
        if (tri[p] == 1) Q8[(size_t)row * n_pad + colidx[p]] = 1;
}

__global__ void __launch_bounds__(256) tc_closing_kernel(const int32_t* __restrict__ rowptr,
                                                         const int32_t* __restrict__ colidx, int n, int64_t nnz,
                                                         const int32_t* __restrict__ tri, const int32_t* __restrict__ t1,
                                                         int32_t* __restrict__ sharp_out, int32_t* __restrict__ lam_out,
                                                         double* __restrict__ c64_out, float* __restrict__ c32_out) {
    // one thread per directed entry (hub rows would serialise a warp-per-row mapping)
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int a = 0, b = n;                       // row of entry p: rowptr[a] <= p < rowptr[b]
    while (b - a > 1) {
        const int m = (a + b) >> 1;
        if ((int64_t)rowptr[m] <= p) a = m; else b = m;
    }
    const int i = a, j = colidx[p];
    const int di = rowptr[i + 1] - rowptr[i];
    const int sj = rowptr[j], dj = rowptr[j + 1] - sj;
    const int q = find_sorted(colidx, sj, dj, i);                // the mirrored entry (j -> i)
    const int dmax = max(di, dj), dmin = min(di, dj);
    const int sharp = di + dj - t1[p] - t1[q];
    const Closing c = closing_value(dmax, dmin, tri[p], 1, sharp, dmax);
    if (sharp_out) sharp_out[p] = sharp;
    if (lam_out) lam_out[p] = dmax;
    if (c64_out) c64_out[p] = c.c64;
    c32_out[p] = c.c32;
}
}  // namespace dcr

extern "C" int64_t dcr_bfc_cuda_flavour_tc_workspace_bytes(int n, int64_t nnz) {
    return (int64_t)tc_layout(nullptr, n, nnz, true).total;
}

extern "C" int dcr_bfc_cuda_flavour_tc(const int32_t* rowptr, const int32_t* colidx, int n, int64_t nnz, int32_t* tri,
                                       int32_t* sharp, int32_t* lam, double* c64, float* c32, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
    if (n <= 0 || nnz <= 0) return 0;
    if (n > 32768) { set_error("dcr_bfc_cuda_flavour_tc: dense tensor-core path is for n <= 32768 (got %d)", n); return 1; }
    if (!tri || !c32) { set_error("dcr_bfc_cuda_flavour_tc: tri and c32 must not be NULL"); return 1; }
    if (workspace_bytes < dcr_bfc_cuda_flavour_tc_workspace_bytes(n, nnz)) {
        set_error("dcr_bfc_cuda_flavour_tc: workspace too small");
        return 1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int np = tc_pad(n);
    const TcWorkspace w = tc_layout(workspace, n, nnz, true);
    const unsigned row_grid = (unsigned)(((int64_t)n * 32 + 255) / 256);
    DCR_CUDA(cudaMemsetAsync(w.A8, 0, 2 * (size_t)np * np, st));     // A8 and Q8 are adjacent
    tc_fill_kernel<<<row_grid, 256, 0, st>>>(rowptr, colidx, n, np, w.A8, w.blkpre);
    DCR_LAUNCH_CHECK();
    CUtensorMap map_a, map_q;
    if (make_tmap(&map_a, w.A8, np) || make_tmap(&map_q, w.Q8, np)) return 1;
    // A·A is symmetric: upper-triangular tiles (half the MMAs), then the mirror pass
    if (launch_edge_gemm(map_a, map_a, w.A8, rowptr, w.blkpre, n, np, tri, true, st)) return 1;
    if (launch_mirror(rowptr, colidx, n, nnz, tri, st)) return 1;
    tc_fill_q_kernel<<<row_grid, 256, 0, st>>>(rowptr, colidx, tri, n, np, w.Q8);
    DCR_LAUNCH_CHECK();
    // Q·A is not ((Q·A)^T = A·Q): every tile
    if (launch_edge_gemm(map_q, map_a, w.A8, rowptr, w.blkpre, n, np, w.t1, false, st)) return 1;
    tc_closing_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(rowptr, colidx, n, nnz, tri, w.t1, sharp, lam, c64, c32);
    DCR_LAUNCH_CHECK();
    return 0;
}
