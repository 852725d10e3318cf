// dcr_bfc_paper.cu — paper-flavour Balanced Forman curvature over a sorted CSR (the headline kernel).
//
// Takes over bfc_edge / bfc (curvature/bfc_naive.py:7-40, :43-52).  Per undirected edge (i,j), with
//   M_i = N(i) \ N(j) \ {j},  M_j = N(j) \ N(i) \ {i}            (the "pure" neighbours of each endpoint)
// the reference's sets are
//   triangles = N(i) ∩ N(j)                                                   (:25)
//   squares_1 = {k in M_i : N(k) ∩ M_j != {}},  squares_2 mirrored            (:26-29)
//   gamma     = max over squares of |N(k) ∩ M_j| resp. |N(m) ∩ M_i|           (:36-37; the "-1" removes i / j)
// i.e. everything is a function of cnt(m) = |N(m) ∩ M_i| for m in M_j and cnt(k) = |N(k) ∩ M_j| for k in M_i.
//
// Mapping to the GPU.  One TEAM per edge — a warp for edges with d_i + d_j <= 512, a whole CTA above that:
//   1. the team builds ONE open-addressing hash table in shared memory holding N(a) ∪ N(b) \ {a,b}, each key
//      tagged 1 (only in N(a)), 2 (only in N(b)) or 3 (both = triangle node); #tag-3 keys = #triangles.
//      (a,b) = (i,j) or (j,i): b is the endpoint whose 2-hop lists are CHEAPER to stream (smaller S_b - d_a);
//   2. ONE scan: for every m with tag 2 the neighbour list N(m) is streamed.  Every match (k in N(m) with tag 1)
//      is an edge of the bipartite graph M_a–M_b, and it is counted TWICE: in a per-list counter (-> cnt(m)) and
//      in a 16-bit counter attached to k's table slot (-> cnt(k)); so the expensive side is never streamed;
//   3. a sweep over the table slots turns the slot counters into (#squares, max) of endpoint a, the per-list
//      counters give those of endpoint b.
// The lists of 32 heads at a time are processed as ONE flat stream (prefix sums of the list lengths in shared
// memory, lane f handles flat element f): coalesced where lists are long, and no idle lanes where they are short
// — half of the 2-hop lists of a power-law graph have fewer than 20 entries.
// Global traffic is at most the algorithmic gather of SURVEY.md §8d (which charges BOTH sides' 2-hop lists): the
// two endpoint lists once and the cheaper side's 2-hop lists once per edge.  The CSR of the benchmark graphs is
// L2-resident, so the kernel is bound by instruction issue / L1 gathers / shared-memory probes, not by DRAM.
// Edges are bucketed by class and by log2(work) on the device (heavy first) and teams pull edges from a global
// counter, so power-law hubs do not serialise the tail.  Grids are persistent: a multiple of the SM count.
#include <algorithm>

#include "dcr_common.cuh"

namespace dcr {

constexpr uint32_t EMPTY = 0xffffffffu;
constexpr uint32_t KEYMASK = 0x3fffffffu;
// Edge classes by the degree d_a of the TESTED endpoint (the one whose neighbour set goes into the hash table):
//   class 0  d_a <= 128    warp team,            512-slot table per warp  (load factor <= 1/4), 48 warps per SM
//   class 1  d_a <= 1024   128-thread CTA team,  2048-slot table          (load factor <= 1/2), 8 CTAs per SM
//   class 2  d_a <= 16384  1024-thread CTA team, 32768-slot table         (load factor <= 1/2), 1 CTA per SM
//   class 3  larger        1024-thread CTA team, table in global memory (L2)
constexpr int N_CLASSES = 4;
constexpr int CLASS_DA0 = 128, CLASS_DA1 = 1024, CLASS_DA2 = 16384;
#ifndef DCR_WARP_SLOTS
#define DCR_WARP_SLOTS 512
#endif
#ifndef DCR_WARP_CTAS
#define DCR_WARP_CTAS 6
#endif
#ifndef DCR_MID_SLOTS
#define DCR_MID_SLOTS 2048
#endif
#ifndef DCR_MID_CTAS
#define DCR_MID_CTAS 8
#endif
constexpr int WARP_SLOTS = DCR_WARP_SLOTS, WARP_CTAS_PER_SM = DCR_WARP_CTAS;
constexpr int WARP_TEAM_WARPS = 8;           // warps (teams) per CTA in the warp-team kernel
constexpr int MID_SLOTS = DCR_MID_SLOTS, MID_THREADS = 128, MID_CTAS_PER_SM = DCR_MID_CTAS;
constexpr int BIG_SLOTS = 32768, BIG_THREADS = 1024;
__host__ __device__ constexpr int heads_per_thread(int team) { return team >= 1024 ? 1 : 4; }   // CTA stream state must fit beside the table
// Membership pre-filter: a hashed bitmap of N(va) (bit index = node id mod B).  Almost every streamed element is NOT
// a neighbour of va, and the bitmap says so with one shared-memory load and no loop; only the few elements whose bit
// is set (true members + d_a/B false positives) go on to the exact hash probe.
constexpr int WARP_BITS = 4096, MID_BITS = 32768, BIG_BITS = 131072, GLOBAL_BITS = 1048576;
__host__ __device__ constexpr int filter_bits(int team, bool global_table) {
    return global_table ? GLOBAL_BITS : (team == 32 ? WARP_BITS : (team >= 1024 ? BIG_BITS : MID_BITS));
}
constexpr int STREAM_INTS = 100;             // per-warp flat-stream state: pre[33] + beg[32] + cnt[32] (+pad)
constexpr int BUCKETS_PER_CLASS = 48;
constexpr int UNROLL = 4;

// Ordering of the work.  Class 0 (warp teams): bucketed by log2(work), heavy first.  Classes 1-3 (CTA teams):
// grouped by the tested endpoint va (counting sort over vertex ids), so that a CTA meets runs of edges with the
// same va and re-uses the hash table of N(va) instead of rebuilding it per edge.
constexpr int BUCKET_TRIVIAL = 255, BUCKET_GROUPED = 200, BUCKET_EXCEPTION = 204;
struct PaperPlan {           // lives at the head of the scratch buffer
    unsigned int hist[BUCKETS_PER_CLASS];        // class-0 buckets
    unsigned int cursor[BUCKETS_PER_CLASS];
    unsigned int bucket_off[BUCKETS_PER_CLASS + 1];
    unsigned int grouped[N_CLASSES + 1];         // [c] = #edges of class c (1..3) grouped by va, [4] = exceptions
    unsigned int exc_cursor;
    unsigned int group_cursor[N_CLASSES];        // range reservation inside classes 1-3
    unsigned int class_begin[N_CLASSES + 1];
    unsigned int next[N_CLASSES];                // work-stealing counters
    unsigned int pad[2];
};

struct PaperArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const int32_t* esrc;
    const int32_t* edst;
    int64_t e_first, e_stride, count;  // this call handles edges e = e_first + t*e_stride, t in [0,count)
    int32_t* out_tri;
    int32_t* out_sq_i;
    int32_t* out_sq_j;
    int32_t* out_gamma;
    double* out_bfc;
    PaperPlan* plan;
    const int64_t* node_s;
    uint8_t* bucket;       // [count]
    uint32_t* order;       // [count] local indices t: class 0 by bucket, classes 1-3 by tested endpoint
    uint32_t* va_cnt;      // [n] #grouped edges whose tested endpoint is v; after the scan: first position in `order`
    uint32_t* va_cur;      // [n] fill cursors
    int n;
    uint32_t* gtables;     // class-2 tables in global memory, gslots per CTA
    uint32_t gslots;
};

__device__ __forceinline__ uint32_t hash_slot(uint32_t key, int shift) { return (key * 2654435761u) >> shift; }

// slot of `key` in the table, or -1.  A table in GLOBAL memory is filled with L2 atomics by other warps of the
// CTA, so it is read with ld.global.cg (L2) — an L1 line fetched during the build phase may be stale.
template <bool GLOBAL>
__device__ __forceinline__ int probe_slot(const uint32_t* tab, uint32_t mask, int shift, uint32_t key, uint32_t& val) {
    uint32_t h = hash_slot(key, shift);
    while (true) {
        const uint32_t v = GLOBAL ? __ldcg(tab + h) : tab[h];
        if (v == EMPTY) return -1;
        if ((v & KEYMASK) == key) { val = v; return (int)h; }
        h = (h + 1) & mask;
    }
}

// insert key with `tag`, or OR the tag into an existing entry; returns true if the key was already present
template <bool GLOBAL>
__device__ __forceinline__ bool insert_or_tag(uint32_t* tab, uint32_t mask, int shift, uint32_t key, uint32_t tag) {
    uint32_t h = hash_slot(key, shift);
    const uint32_t val = key | (tag << 30);
    while (true) {
        uint32_t v = GLOBAL ? __ldcg(tab + h) : tab[h];
        if (v == EMPTY) {
            v = atomicCAS(&tab[h], EMPTY, val);
            if (v == EMPTY) return false;
        }
        if ((v & KEYMASK) == key) {
            atomicOr(&tab[h], tag << 30);
            return true;
        }
        h = (h + 1) & mask;
    }
}

// bfc_naive.py:31-32 / :39-40 evaluated left to right in fp64 with explicitly rounded operations
__device__ __forceinline__ double paper_value(int d1, int d2, int tri, int sq1, int sq2, int gamma) {
    const int dmax = max(d1, d2), dmin = min(d1, d2);
    double t = __ddiv_rn(2.0, (double)d1);
    t = __dadd_rn(t, __ddiv_rn(2.0, (double)d2));
    t = __dadd_rn(t, -2.0);
    t = __dadd_rn(t, __ddiv_rn((double)(2 * (long long)tri), (double)dmax));
    t = __dadd_rn(t, __ddiv_rn((double)tri, (double)dmin));
    if (sq1 > 0 && sq2 > 0) {
        double u = __ddiv_rn(1.0, (double)gamma);
        u = __ddiv_rn(u, (double)dmax);
        u = __dmul_rn(u, (double)(sq1 + sq2));
        t = __dadd_rn(t, u);
    }
    return t;
}

// ------------------------------------------------------------------------------------------------------------
// planning kernels: S_v, per-edge class/bucket, bucket offsets, order
// ------------------------------------------------------------------------------------------------------------
// S_v = sum of the degrees of v's neighbours, via a degree array (one 4-byte gather per entry instead of two
// dependent row-offset loads).
__global__ void degree_kernel(const int32_t* __restrict__ rowptr, int n, int32_t* __restrict__ deg) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) deg[v] = rowptr[v + 1] - rowptr[v];
}
__global__ void node_s_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ colidx,
                              const int32_t* __restrict__ deg, int n, int64_t* __restrict__ node_s) {
    const int lane = threadIdx.x & 31;
    const int v = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (v >= n) return;
    const int b = rowptr[v], e = rowptr[v + 1];
    int64_t s = 0;
    for (int p = b + lane; p < e; p += 32) s += deg[colidx[p]];
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
    if (lane == 0) node_s[v] = s;
}

__device__ __forceinline__ int class_of_degree(int da) {
    return da <= CLASS_DA0 ? 0 : (da <= CLASS_DA1 ? 1 : (da <= CLASS_DA2 ? 2 : 3));
}

__global__ void __launch_bounds__(256) classify_kernel(PaperArgs a) {
    __shared__ unsigned int s_hist[BUCKETS_PER_CLASS];
    __shared__ unsigned int s_grp[N_CLASSES + 1];
    for (int b = threadIdx.x; b < BUCKETS_PER_CLASS; b += blockDim.x) s_hist[b] = 0;
    if (threadIdx.x <= N_CLASSES) s_grp[threadIdx.x] = 0;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < a.count) {
        const int64_t e = a.e_first + t * a.e_stride;
        const int i = a.esrc[e], j = a.edst[e];
        const int di = a.rowptr[i + 1] - a.rowptr[i], dj = a.rowptr[j + 1] - a.rowptr[j];
        if (min(di, dj) <= 1) {  // bfc_naive.py:18-19: deg_min == 1 -> 0 (no triangle is possible either)
            a.out_tri[t] = 0; a.out_sq_i[t] = 0; a.out_sq_j[t] = 0; a.out_gamma[t] = 0; a.out_bfc[t] = 0.0;
            a.bucket[t] = BUCKET_TRIVIAL;
        } else {
            const long long ca = a.node_s[j] - di, cb = a.node_s[i] - dj;     // 2-hop entries behind j / behind i
            const bool swapped = cb < ca;
            const int va = swapped ? j : i;                                    // tested endpoint (see the edge kernel)
            const int da = swapped ? dj : di, db = swapped ? di : dj;
            const int cls = class_of_degree(da);
            if (db > 65535) {            // the 16-bit slot counters of the shared-memory tables count up to d_b lists
                a.bucket[t] = BUCKET_EXCEPTION;
                atomicAdd(&s_grp[N_CLASSES], 1u);
            } else if (cls == 0) {
                const long long work = min(ca, cb) + 4LL * (di + dj);
                const int lg = 63 - __clzll(work | 1);
                const int b = BUCKETS_PER_CLASS - 1 - min(lg, BUCKETS_PER_CLASS - 1);   // heavy first
                a.bucket[t] = (uint8_t)b;
                atomicAdd(&s_hist[b], 1u);
            } else {
                a.bucket[t] = (uint8_t)(BUCKET_GROUPED + cls);
                atomicAdd(&s_grp[cls], 1u);
                atomicAdd(&a.va_cnt[va], 1u);
            }
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < BUCKETS_PER_CLASS; b += blockDim.x)
        if (s_hist[b]) atomicAdd(&a.plan->hist[b], s_hist[b]);
    if (threadIdx.x <= N_CLASSES && s_grp[threadIdx.x]) atomicAdd(&a.plan->grouped[threadIdx.x], s_grp[threadIdx.x]);
}

// Bucket offsets of class 0 and the class ranges (one thread: 48 + 4 values).
__global__ void plan_ranges_kernel(PaperArgs a) {
    PaperPlan* plan = a.plan;
    if (threadIdx.x == 0) {
        unsigned int acc = 0;
        plan->class_begin[0] = 0;
        for (int b = 0; b < BUCKETS_PER_CLASS; ++b) { plan->bucket_off[b] = acc; acc += plan->hist[b]; }
        plan->bucket_off[BUCKETS_PER_CLASS] = acc;
        for (int c = 1; c < N_CLASSES; ++c) { plan->class_begin[c] = acc; acc += plan->grouped[c]; }
        acc += plan->grouped[N_CLASSES];               // exceptions close class 3
        plan->class_begin[N_CLASSES] = acc;
    }
}

// Every vertex that is the tested endpoint of grouped edges reserves a contiguous range of `order` inside its
// class (the class is a function of the vertex degree).  The order of the ranges is irrelevant — only contiguity
// matters for the table reuse — so one atomicAdd per such vertex replaces a scan over all vertices.
__global__ void __launch_bounds__(256) plan_groups_kernel(PaperArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.n) return;
    const unsigned int c = a.va_cnt[v];
    if (c == 0) return;
    const int cls = class_of_degree(a.rowptr[v + 1] - a.rowptr[v]);
    a.va_cnt[v] = a.plan->class_begin[cls] + atomicAdd(&a.plan->group_cursor[cls], c);
}

__global__ void __launch_bounds__(256) order_kernel(PaperArgs a) {
    // class 0, block-aggregated: one global atomic per (block, bucket) reserves a range, ranks come from shared memory
    __shared__ unsigned int s_cnt[BUCKETS_PER_CLASS];
    __shared__ unsigned int s_base[BUCKETS_PER_CLASS];
    for (int b = threadIdx.x; b < BUCKETS_PER_CLASS; b += blockDim.x) s_cnt[b] = 0;
    __syncthreads();
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int b = BUCKET_TRIVIAL;
    unsigned int r = 0;
    if (t < a.count) {
        b = a.bucket[t];
        if (b < BUCKETS_PER_CLASS) {
            r = atomicAdd(&s_cnt[b], 1u);
        } else if (b == BUCKET_EXCEPTION) {
            const unsigned int pos = a.plan->class_begin[N_CLASSES] - a.plan->grouped[N_CLASSES] +
                                     atomicAdd(&a.plan->exc_cursor, 1u);
            a.order[pos] = (uint32_t)t;
        } else if (b != BUCKET_TRIVIAL) {          // grouped by tested endpoint
            const int64_t e = a.e_first + t * a.e_stride;
            const int i = a.esrc[e], j = a.edst[e];
            const int di = a.rowptr[i + 1] - a.rowptr[i], dj = a.rowptr[j + 1] - a.rowptr[j];
            const int va = (a.node_s[i] - dj) < (a.node_s[j] - di) ? j : i;
            a.order[a.va_cnt[va] + atomicAdd(&a.va_cur[va], 1u)] = (uint32_t)t;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q < BUCKETS_PER_CLASS; q += blockDim.x)
        if (s_cnt[q]) s_base[q] = a.plan->bucket_off[q] + atomicAdd(&a.plan->cursor[q], s_cnt[q]);
    __syncthreads();
    if (b < BUCKETS_PER_CLASS) a.order[s_base[b] + r] = (uint32_t)t;
}

template <bool CTA_TEAM>
__device__ __forceinline__ void team_sync() {
    if (CTA_TEAM) __syncthreads(); else __syncwarp();
}

// Slot counters: 16 bit per slot packed in 32-bit words for the shared-memory tables (a count is at most the number
// of streamed lists, and classes 0-2 cap it far below 65536 only through d_b — see the saturation note in the
// kernel), 32 bit per slot for the global-memory tables.
template <bool GLOBAL>
__device__ __forceinline__ void slot_count_add(uint32_t* cnt, int h) {
    if (GLOBAL) atomicAdd(&cnt[h], 1u);
    else atomicAdd(&cnt[h >> 1], (h & 1) ? 0x10000u : 1u);
}
template <bool GLOBAL>
__device__ __forceinline__ uint32_t slot_count_get(const uint32_t* cnt, uint32_t h) {
    if (GLOBAL) return __ldcg(cnt + h);
    return (cnt[h >> 1] >> ((h & 1) * 16)) & 0xffffu;
}

// The flat stream: `pre` = exclusive prefix sums of the list lengths (pre[l+1]-pre[l] = length of list l), `beg` =
// first CSR slot of each list, `lcnt` = per-list match counters, all in shared memory.  This warp handles the flat
// elements [f_begin, f_end) in windows of 32*UNROLL; lane f%32 takes element f.  `lw` = a list index with
// pre[lw] <= f_begin.  Windows that lie inside one list (most elements belong to long lists) take a fast path with
// no per-element owner search.  An element first meets the bitmap filter `bm`; only if its bit is set the exact
// probe runs.  A match (key present with tag 1, and not vb itself — the table holds all of N(va)) bumps the list's
// counter and the matched key's slot counter.
template <bool GLOBAL>
__device__ __forceinline__ void flat_scan(const int32_t* __restrict__ colidx, const uint32_t* tab, uint32_t* cnt,
                                          uint32_t mask, int shift, const uint32_t* bm, uint32_t bmask, const int* pre,
                                          const int* beg, int* lcnt, int lw, int f_begin, int f_end, int lane, int vb) {
    for (int F = f_begin; F < f_end; F += 32 * UNROLL) {
        while (pre[lw + 1] <= F) ++lw;                       // warp-uniform
        const int win = min(32 * UNROLL, f_end - F);
        int k[UNROLL], lu[UNROLL];
        if (pre[lw + 1] - F >= win) {                        // the whole window lies inside list lw
            const int32_t* src = colidx + beg[lw] + (F - pre[lw]);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int idx = 32 * u + lane;
                k[u] = idx < win ? src[idx] : -1;
                lu[u] = lw;
            }
        } else {
            int l = lw;
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int f = F + 32 * u + lane;
                k[u] = -1;
                lu[u] = 0;
                if (f < f_end) {
                    while (pre[l + 1] <= f) ++l;             // monotone in f: amortised O(1)
                    lu[u] = l;
                    k[u] = colidx[beg[l] + (f - pre[l])];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            // vb itself sits in every streamed list (m in N(vb)) and in the table (vb in N(va)): dropping it BEFORE the
            // filter keeps the exact probe off the common path
            const uint32_t kk = (uint32_t)k[u];
            if (k[u] >= 0 && k[u] != vb && ((bm[(kk & bmask) >> 5] >> (kk & 31u)) & 1u)) {
                uint32_t v;
                const int h = probe_slot<GLOBAL>(tab, mask, shift, kk, v);
                if (h >= 0 && (v >> 30) == 1u) {
                    atomicAdd(&lcnt[lu[u]], 1);
                    slot_count_add<GLOBAL>(cnt, h);
                }
            }
        }
    }
}

// (begin, length) of the neighbour list of head m if m is a PURE neighbour of vb (not va, not in the table), else (0,0)
template <bool GLOBAL>
__device__ __forceinline__ void pure_head(const PaperArgs& a, const uint32_t* tab, uint32_t mask, int shift, int m,
                                          int va, int& b, int& d) {
    b = 0;
    d = 0;
    uint32_t v;
    if (m != va && probe_slot<GLOBAL>(tab, mask, shift, (uint32_t)m, v) < 0) {
        b = a.rowptr[m];
        d = a.rowptr[m + 1] - b;
    }
}

// Warp team: the lists of the pure heads among list[c0 .. c0+32) as one flat stream.  `st` = this warp's stream
// state.  Accumulates (#lists with a match, largest per-list count) per lane.
template <bool GLOBAL>
__device__ __forceinline__ void scan_chunk(const PaperArgs& a, const uint32_t* tab, uint32_t* cnt, uint32_t mask,
                                           int shift, const uint32_t* bm, uint32_t bmask, int list_begin, int list_len,
                                           int c0, int va, int vb, int* st, int lane, int& sq, int& gmax) {
    int* pre = st;            // [33]
    int* beg = st + 33;       // [32]
    int* lcnt = st + 65;      // [32]
    const int t = c0 + lane;
    int mb = 0, md = 0;
    if (t < list_len) pure_head<GLOBAL>(a, tab, mask, shift, a.colidx[list_begin + t], va, mb, md);
    int inc = md;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int up = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += up;
    }
    const int total = __shfl_sync(FULL, inc, 31);
    if (total == 0) return;
    pre[lane + 1] = inc;
    if (lane == 0) pre[0] = 0;
    beg[lane] = mb;
    lcnt[lane] = 0;
    __syncwarp();
    flat_scan<GLOBAL>(a.colidx, tab, cnt, mask, shift, bm, bmask, pre, beg, lcnt, 0, 0, total, lane, vb);
    __syncwarp();
    const int c = lcnt[lane];
    sq += c > 0;
    gmax = max(gmax, c);
    __syncwarp();
}

// CTA team: up to HEADS = heads_per_thread*THREADS heads per round; the CTA-wide flat stream is cut into equal contiguous ranges,
// one per warp, so every warp streams the same number of elements whatever the list-length distribution (a hub's
// list next to twenty short ones does not serialise).  `cs` = pre[HEADS+1] | beg[HEADS] | cnt[HEADS].
template <int THREADS, bool GLOBAL>
__device__ __forceinline__ void scan_cta(const PaperArgs& a, const uint32_t* tab, uint32_t* cnt, uint32_t mask, int shift,
                                         const uint32_t* bm, uint32_t bmask, int list_begin, int list_len, int va,
                                         int vb, int* cs, int* s_warp_tot, int& sq, int& gmax) {
    constexpr int HPT = heads_per_thread(THREADS);   // heads per thread
    constexpr int HEADS = HPT * THREADS;
    constexpr int NW = THREADS / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* pre = cs;                 // [HEADS + 1]
    int* beg = cs + HEADS + 1;     // [HEADS]
    int* lcnt = beg + HEADS;       // [HEADS]
    for (int h0 = 0; h0 < list_len; h0 += HEADS) {
        const int nh = min(HEADS, list_len - h0);
        int deg[HPT], run = 0;
#pragma unroll
        for (int q = 0; q < HPT; ++q) {
            const int t = tid * HPT + q;
            int b = 0, d = 0;
            if (t < nh) pure_head<GLOBAL>(a, tab, mask, shift, a.colidx[list_begin + h0 + t], va, b, d);
            beg[t] = b;
            lcnt[t] = 0;
            deg[q] = d;
            run += d;
        }
        int inc = run;                            // exclusive scan of `run` over the CTA
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += up;
        }
        if (lane == 31) s_warp_tot[warp] = inc;
        __syncthreads();
        int off = inc - run;
        for (int w = 0; w < warp; ++w) off += s_warp_tot[w];
#pragma unroll
        for (int q = 0; q < HPT; ++q) {
            pre[tid * HPT + q] = off;
            off += deg[q];
        }
        if (tid == THREADS - 1) pre[HEADS] = off;
        __syncthreads();
        const int total = pre[HEADS];
        if (total > 0) {
            const int per_warp = ((total + NW * 32 - 1) / (NW * 32)) * 32;
            const int f_begin = warp * per_warp, f_end = min(total, f_begin + per_warp);
            if (f_begin < f_end) {
                int lo = 0, hi = HEADS;            // first l with pre[l+1] > f_begin (warp-uniform search)
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (pre[mid + 1] <= f_begin) lo = mid + 1; else hi = mid;
                }
                flat_scan<GLOBAL>(a.colidx, tab, cnt, mask, shift, bm, bmask, pre, beg, lcnt, lo, f_begin, f_end, lane, vb);
            }
            __syncthreads();
            for (int t = tid; t < nh; t += THREADS) {
                const int c = lcnt[t];
                sq += c > 0;
                gmax = max(gmax, c);
            }
        }
        __syncthreads();
    }
}

// TEAM = threads per team (32 = warp team, several teams per CTA; otherwise the CTA is the team).
template <int TEAM, int MAX_SLOTS, bool GLOBAL_TABLE>
__global__ void __launch_bounds__(TEAM > 32 ? TEAM : WARP_TEAM_WARPS * 32,
                                  TEAM == BIG_THREADS ? 1 : (TEAM == MID_THREADS ? MID_CTAS_PER_SM : WARP_CTAS_PER_SM))
paper_edge_kernel(PaperArgs a, int cls) {
    constexpr bool CTA_TEAM = TEAM > 32;
    constexpr int NWARPS = CTA_TEAM ? TEAM / 32 : WARP_TEAM_WARPS;
    constexpr int CTA_STREAM = 3 * heads_per_thread(TEAM) * TEAM + 8;
    extern __shared__ uint32_t smem_dyn[];
    __shared__ unsigned int s_idx;
    __shared__ int s_warp_tot[NWARPS];
    __shared__ int s_red[5];  // tri, sq(lists), g(lists), sq(slots), g(slots)

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int team_threads = CTA_TEAM ? TEAM : 32;
    const int team_tid = CTA_TEAM ? (int)threadIdx.x : lane;
    // shared-memory carve-up: [stream state (per warp | per CTA)][bitmap filter(s)][table keys][slot counters]
    constexpr int FBITS = filter_bits(TEAM, GLOBAL_TABLE);
    constexpr uint32_t bmask = (uint32_t)FBITS - 1u;
    int* st = (int*)smem_dyn + (CTA_TEAM ? 0 : warp * STREAM_INTS);
    uint32_t* bm = smem_dyn + (CTA_TEAM ? CTA_STREAM : NWARPS * STREAM_INTS + warp * (FBITS / 32));
    uint32_t* sm_tab = smem_dyn + (CTA_TEAM ? CTA_STREAM + FBITS / 32 : NWARPS * (STREAM_INTS + FBITS / 32));
    uint32_t* tab;
    uint32_t* cnt;
    if (GLOBAL_TABLE) {
        tab = a.gtables + (size_t)blockIdx.x * a.gslots * 2;
        cnt = tab + a.gslots;
    } else if (CTA_TEAM) {
        tab = sm_tab;
        cnt = sm_tab + MAX_SLOTS;
    } else {
        tab = sm_tab + warp * (MAX_SLOTS + MAX_SLOTS / 2);
        cnt = tab + MAX_SLOTS;
    }
    const unsigned int cbeg = a.plan->class_begin[cls];
    const unsigned int cnum = a.plan->class_begin[cls + 1] - cbeg;
    const uint32_t max_slots = GLOBAL_TABLE ? a.gslots : (uint32_t)MAX_SLOTS;
    // edges per work-stealing step: a run of one va, usually — but never so long that the CTAs run out of steps
    const unsigned int GRAB = CTA_TEAM ? min(16u, max(1u, cnum / (gridDim.x * 8u))) : 1u;

    int cur_va = -1;                 // vertex whose neighbour set is in the table (CTA teams re-use it across edges)
    uint32_t slots = 0, mask = 0;
    int shift = 0;
    while (true) {
        unsigned int idx0;
        if (CTA_TEAM) {
            if (threadIdx.x == 0) s_idx = atomicAdd(&a.plan->next[cls], GRAB);
            __syncthreads();
            idx0 = s_idx;
        } else {
            idx0 = 0;
            if (lane == 0) idx0 = atomicAdd(&a.plan->next[cls], GRAB);
            idx0 = __shfl_sync(FULL, idx0, 0);
        }
        if (idx0 >= cnum) break;
        for (unsigned int q = 0; q < GRAB && idx0 + q < cnum; ++q) {
            if (CTA_TEAM) {
                if (threadIdx.x == 0) s_red[0] = s_red[1] = s_red[2] = s_red[3] = s_red[4] = 0;
                __syncthreads();
            }
            const uint32_t t = a.order[cbeg + idx0 + q];
            const int64_t e = a.e_first + (int64_t)t * a.e_stride;
            const int i = a.esrc[e], j = a.edst[e];
            const int di = a.rowptr[i + 1] - a.rowptr[i], dj = a.rowptr[j + 1] - a.rowptr[j];
            // (va, vb): vb = endpoint whose 2-hop lists are streamed (the cheaper side), va = the tested side
            const bool swapped = (a.node_s[i] - dj) < (a.node_s[j] - di);
            const int va = swapped ? j : i, vb = swapped ? i : j;
            const int sa = a.rowptr[va], da = swapped ? dj : di;
            const int sb = a.rowptr[vb], db = swapped ? di : dj;

            if (!CTA_TEAM || va != cur_va) {
                // table of N(va), every key with tag 1: power of two >= 8*d_a (load factor <= 1/8 keeps probe
                // chains short and uniform across the lanes of a warp), capped by the class's table
                const int lg = min(32 - __clz(max(8 * da, 64) - 1), 31 - __clz(max_slots));
                slots = 1u << lg;
                mask = slots - 1;
                shift = 32 - lg;
                for (uint32_t s = team_tid; s < slots; s += team_threads) tab[s] = EMPTY;
                for (uint32_t s = team_tid; s < (GLOBAL_TABLE ? slots : slots / 2); s += team_threads) cnt[s] = 0u;
                for (uint32_t s = team_tid; s < (uint32_t)FBITS / 32; s += team_threads) bm[s] = 0u;
                team_sync<CTA_TEAM>();
                for (int p = team_tid; p < da; p += team_threads) {
                    const uint32_t k = (uint32_t)a.colidx[sa + p];
                    insert_or_tag<GLOBAL_TABLE>(tab, mask, shift, k, 1u);
                    atomicOr(&bm[(k & bmask) >> 5], 1u << (k & 31u));
                }
                team_sync<CTA_TEAM>();
                cur_va = va;
            }
            // common neighbours: heads of N(vb) found in the table get bit 31 (tag 3: never a match, never a head)
            int tri = 0;
            for (int p = team_tid; p < db; p += team_threads) {
                const int m = a.colidx[sb + p];
                uint32_t v;
                const int h = (m == va) ? -1 : probe_slot<GLOBAL_TABLE>(tab, mask, shift, (uint32_t)m, v);
                if (h >= 0) {
                    atomicOr(&tab[h], 2u << 30);
                    ++tri;
                }
            }
            tri = warp_sum(tri);
            if (CTA_TEAM) {
                if (lane == 0 && tri) atomicAdd(&s_red[0], tri);
            }
            team_sync<CTA_TEAM>();

            // the scan: lists of the pure neighbours of vb, matches against the pure neighbours of va
            int sqL = 0, gL = 0, sqS = 0, gS = 0;
            if (CTA_TEAM) {
                scan_cta<TEAM, GLOBAL_TABLE>(a, tab, cnt, mask, shift, bm, bmask, sb, db, va, vb, st, s_warp_tot, sqL, gL);
                sqL = warp_sum(sqL);
                gL = warp_max(gL);
                if (lane == 0 && sqL) { atomicAdd(&s_red[1], sqL); atomicMax(&s_red[2], gL); }
                __syncthreads();
                sqL = s_red[1]; gL = s_red[2]; tri = s_red[0];
            } else {
                for (int c0 = 0; c0 < db; c0 += 32)
                    scan_chunk<GLOBAL_TABLE>(a, tab, cnt, mask, shift, bm, bmask, sb, db, c0, va, vb, st, lane, sqL, gL);
                sqL = warp_sum(sqL);
                gL = warp_max(gL);
                __syncwarp();
            }
            // the collect: slot counters of the pure neighbours of va -> squares at va (empty iff the scan found
            // nothing).  Walks N(va) (d_a probes) instead of sweeping the table; CTA teams zero the counters they
            // read so the table is clean for the next edge of the same va.
            if (sqL > 0) {
                for (int p = team_tid; p < da; p += team_threads) {
                    const int k = a.colidx[sa + p];
                    uint32_t v;
                    const int h = probe_slot<GLOBAL_TABLE>(tab, mask, shift, (uint32_t)k, v);
                    if (h >= 0 && (v >> 30) == 1u) {
                        const int c = (int)slot_count_get<GLOBAL_TABLE>(cnt, (uint32_t)h);
                        if (c > 0) {
                            ++sqS;
                            gS = max(gS, c);
                            if (CTA_TEAM) {
                                if (GLOBAL_TABLE) cnt[h] = 0u;
                                else atomicAnd(&cnt[h >> 1], (h & 1) ? 0x0000ffffu : 0xffff0000u);
                            }
                        }
                    }
                }
                sqS = warp_sum(sqS);
                gS = warp_max(gS);
                if (CTA_TEAM) {
                    if (lane == 0 && sqS) { atomicAdd(&s_red[3], sqS); atomicMax(&s_red[4], gS); }
                    __syncthreads();
                    sqS = s_red[3]; gS = s_red[4];
                }
            }
            if (team_tid == 0) {
                const int sq_i = swapped ? sqL : sqS, sq_j = swapped ? sqS : sqL;
                const int gamma = (sqL > 0 && sqS > 0) ? max(gL, gS) : 0;
                a.out_tri[t] = tri;
                a.out_sq_i[t] = sq_i;
                a.out_sq_j[t] = sq_j;
                a.out_gamma[t] = gamma;      // the fp64 value is computed by paper_value_kernel (one thread per edge)
            }
            if (CTA_TEAM) {
                // undo the common-neighbour marks so the table is N(va) with tag 1 again
                if (tri > 0) {
                    for (int p = team_tid; p < db; p += team_threads) {
                        const int m = a.colidx[sb + p];
                        uint32_t v;
                        const int h = (m == va) ? -1 : probe_slot<GLOBAL_TABLE>(tab, mask, shift, (uint32_t)m, v);
                        if (h >= 0) atomicAnd(&tab[h], 0x7fffffffu);
                    }
                }
            }
            team_sync<CTA_TEAM>();  // table / s_red reuse
        }
    }
}

// bfc_naive.py:31-32 / :39-40 for every edge of the call, one thread per edge: the eight fp64 divisions would
// otherwise run on one lane of the edge's team.
__global__ void paper_value_kernel(PaperArgs a) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.count) return;
    const int64_t e = a.e_first + t * a.e_stride;
    const int i = a.esrc[e], j = a.edst[e];
    const int di = a.rowptr[i + 1] - a.rowptr[i], dj = a.rowptr[j + 1] - a.rowptr[j];
    if (min(di, dj) <= 1) return;      // written by classify_kernel: the int 0 of bfc_naive.py:18-19
    a.out_bfc[t] = paper_value(di, dj, a.out_tri[t], a.out_sq_i[t], a.out_sq_j[t], a.out_gamma[t]);
}

}  // namespace dcr

using namespace dcr;

static inline uint32_t next_pow2_u32(uint64_t v) {
    uint32_t p = 64;
    while (p < v) p <<= 1;
    return p;
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ScratchLayout {
    size_t plan, node_s, bucket, order, va_cnt, va_cur, gtables, total;
    uint32_t gslots;
    int g_ctas;
};

static ScratchLayout scratch_layout(int n, int max_degree, int64_t count) {
    ScratchLayout L;
    size_t off = 0;
    L.plan = off; off = align_up(off + sizeof(PaperPlan), 256);
    L.node_s = off; off = align_up(off + (size_t)n * sizeof(int64_t), 256);
    L.bucket = off; off = align_up(off + (size_t)count, 256);
    L.order = off; off = align_up(off + (size_t)count * sizeof(uint32_t), 256);
    L.va_cnt = off; off = align_up(off + (size_t)n * sizeof(uint32_t), 256);
    L.va_cur = off; off = align_up(off + (size_t)n * sizeof(uint32_t), 256);
    L.gslots = 0;
    L.g_ctas = 0;
    L.gtables = off;
    if (max_degree > CLASS_DA2) {  // some edge may need a table beyond shared memory (or 32-bit slot counters)
        L.gslots = next_pow2_u32((uint64_t)max_degree * 4);
        L.g_ctas = sm_count();
        off = align_up(off + (size_t)L.g_ctas * L.gslots * 2 * sizeof(uint32_t), 256);   // keys + counters
    }
    L.total = off;
    return L;
}

extern "C" int64_t dcr_bfc_paper_scratch_bytes(int n, int max_degree, int64_t count) {
    return (int64_t)scratch_layout(n, max_degree, count).total;
}

extern "C" int dcr_bfc_paper(const int32_t* rowptr, const int32_t* colidx, int n, int max_degree,
                             const int32_t* esrc, const int32_t* edst, int64_t e_first, int64_t e_stride,
                             int64_t count, int32_t* out_tri, int32_t* out_sq_i, int32_t* out_sq_j,
                             int32_t* out_gamma, double* out_bfc, void* scratch, int64_t scratch_bytes,
                             void* ev_edge_begin, void* ev_edge_end, void* stream) {
    if (count <= 0 || n <= 0) return 0;
    if (count > 0xfffffff0LL) { set_error("dcr_bfc_paper: more than 2^32 edges per call"); return 1; }
    if (n >= (1 << 30) - 1) { set_error("dcr_bfc_paper: node ids must fit 30 bits (table keys carry 2 tag bits)"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const ScratchLayout L = scratch_layout(n, max_degree, count);
    if ((int64_t)L.total > scratch_bytes) {
        set_error("dcr_bfc_paper: scratch too small (%lld < %lld bytes)", (long long)scratch_bytes, (long long)L.total);
        return 1;
    }
    char* base = (char*)scratch;
    PaperArgs a;
    a.rowptr = rowptr; a.colidx = colidx; a.esrc = esrc; a.edst = edst;
    a.e_first = e_first; a.e_stride = e_stride; a.count = count;
    a.out_tri = out_tri; a.out_sq_i = out_sq_i; a.out_sq_j = out_sq_j; a.out_gamma = out_gamma; a.out_bfc = out_bfc;
    a.plan = (PaperPlan*)(base + L.plan);
    int64_t* node_s = (int64_t*)(base + L.node_s);
    a.node_s = node_s;
    a.bucket = (uint8_t*)(base + L.bucket);
    a.order = (uint32_t*)(base + L.order);
    a.va_cnt = (uint32_t*)(base + L.va_cnt);
    a.va_cur = (uint32_t*)(base + L.va_cur);
    a.n = n;
    a.gtables = (uint32_t*)(base + L.gtables);
    a.gslots = L.gslots;

    DCR_CUDA(cudaMemsetAsync(a.plan, 0, sizeof(PaperPlan), st));
    DCR_CUDA(cudaMemsetAsync(a.va_cnt, 0, (size_t)n * sizeof(uint32_t), st));
    int32_t* deg = (int32_t*)a.va_cur;     // va_cur is not used before order_kernel; cleared again below
    degree_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rowptr, n, deg);
    DCR_LAUNCH_CHECK();
    node_s_kernel<<<(unsigned)(((int64_t)n * 32 + 255) / 256), 256, 0, st>>>(rowptr, colidx, deg, n, node_s);
    DCR_LAUNCH_CHECK();
    DCR_CUDA(cudaMemsetAsync(a.va_cur, 0, (size_t)n * sizeof(uint32_t), st));
    const unsigned tb = (unsigned)((count + 255) / 256);
    classify_kernel<<<tb, 256, 0, st>>>(a);
    DCR_LAUNCH_CHECK();
    plan_ranges_kernel<<<1, 32, 0, st>>>(a);
    DCR_LAUNCH_CHECK();
    plan_groups_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a);
    DCR_LAUNCH_CHECK();
    order_kernel<<<tb, 256, 0, st>>>(a);
    DCR_LAUNCH_CHECK();

    const int sms = sm_count();
    if (ev_edge_begin) DCR_CUDA(cudaEventRecord((cudaEvent_t)ev_edge_begin, st));
    // heavy classes first: they own the long tail.  Persistent grids = SM count x resident CTAs per SM.
    static bool attr_done_dev[MAX_DEVICES] = {false};
    const int dev = current_device();
    bool& attr_done = attr_done_dev[dev];
    constexpr int big_stream = 3 * heads_per_thread(BIG_THREADS) * BIG_THREADS + 8;
    constexpr int mid_stream = 3 * heads_per_thread(MID_THREADS) * MID_THREADS + 8;
    const int smem_x = (big_stream + GLOBAL_BITS / 32) * (int)sizeof(int);
    const int smem_big = (big_stream + BIG_BITS / 32 + BIG_SLOTS + BIG_SLOTS / 2) * (int)sizeof(uint32_t);
    const int smem_mid = (mid_stream + MID_BITS / 32 + MID_SLOTS + MID_SLOTS / 2) * (int)sizeof(uint32_t);
    const int smem_warp = WARP_TEAM_WARPS * (STREAM_INTS + WARP_BITS / 32 + WARP_SLOTS + WARP_SLOTS / 2) * (int)sizeof(uint32_t);
    if (!attr_done) {
        DCR_CUDA(cudaFuncSetAttribute(paper_edge_kernel<BIG_THREADS, BIG_SLOTS, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem_big));
        DCR_CUDA(cudaFuncSetAttribute(paper_edge_kernel<MID_THREADS, MID_SLOTS, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem_mid));
        DCR_CUDA(cudaFuncSetAttribute(paper_edge_kernel<32, WARP_SLOTS, false>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem_warp));
        DCR_CUDA(cudaFuncSetAttribute(paper_edge_kernel<BIG_THREADS, BIG_SLOTS, true>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, smem_x));
        attr_done = true;
    }
    // The class kernels are independent (disjoint edges, disjoint outputs).  They are launched on three streams
    // forked from `st` — heaviest class first — so that, as the persistent CTAs of one class run out of work, CTAs
    // of the next class take over the freed SMs instead of waiting for the slowest CTA (tail filling).
    static cudaStream_t aux_dev[MAX_DEVICES][2] = {};
    static cudaEvent_t fork_dev[MAX_DEVICES] = {}, join_dev[MAX_DEVICES][2] = {};
    cudaStream_t* aux = aux_dev[dev];
    cudaEvent_t& ev_fork = fork_dev[dev];
    cudaEvent_t* ev_join = join_dev[dev];
    if (!aux[0]) {
        for (int q = 0; q < 2; ++q) {
            DCR_CUDA(cudaStreamCreateWithFlags(&aux[q], cudaStreamNonBlocking));
            DCR_CUDA(cudaEventCreateWithFlags(&ev_join[q], cudaEventDisableTiming));
        }
        DCR_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
    }
    DCR_CUDA(cudaEventRecord(ev_fork, st));
    DCR_CUDA(cudaStreamWaitEvent(aux[0], ev_fork, 0));
    DCR_CUDA(cudaStreamWaitEvent(aux[1], ev_fork, 0));
    if (L.gslots) {
        paper_edge_kernel<BIG_THREADS, BIG_SLOTS, true><<<L.g_ctas, BIG_THREADS, smem_x, st>>>(a, 3);
        DCR_LAUNCH_CHECK();
    }
    paper_edge_kernel<BIG_THREADS, BIG_SLOTS, false><<<sms, BIG_THREADS, smem_big, st>>>(a, 2);
    DCR_LAUNCH_CHECK();
    paper_edge_kernel<MID_THREADS, MID_SLOTS, false><<<sms * MID_CTAS_PER_SM, MID_THREADS, smem_mid, aux[0]>>>(a, 1);
    DCR_LAUNCH_CHECK();
    paper_edge_kernel<32, WARP_SLOTS, false><<<sms * WARP_CTAS_PER_SM, WARP_TEAM_WARPS * 32, smem_warp, aux[1]>>>(a, 0);
    DCR_LAUNCH_CHECK();
    for (int q = 0; q < 2; ++q) {
        DCR_CUDA(cudaEventRecord(ev_join[q], aux[q]));
        DCR_CUDA(cudaStreamWaitEvent(st, ev_join[q], 0));
    }
    paper_value_kernel<<<tb, 256, 0, st>>>(a);
    DCR_LAUNCH_CHECK();
    if (ev_edge_end) DCR_CUDA(cudaEventRecord((cudaEvent_t)ev_edge_end, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// multi-GPU: re-interleave the all-gathered per-rank shards (rank r holds edges e = r + t*world at local t)
// ------------------------------------------------------------------------------------------------------------
namespace dcr {
__global__ void unshard_kernel(const unsigned char* __restrict__ gathered, int world, int64_t chunk, int64_t n_edges,
                               int32_t* __restrict__ tri, int32_t* __restrict__ sq_i, int32_t* __restrict__ sq_j,
                               int32_t* __restrict__ gamma, double* __restrict__ bfc) {
    // per-rank block layout (bytes): bfc[chunk] f64 | tri[chunk] | sq_i[chunk] | sq_j[chunk] | gamma[chunk] i32
    const int64_t block = chunk * 24;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n_edges; e += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(e % world);
        const int64_t t = e / world;
        const unsigned char* base = gathered + r * block;
        bfc[e] = ((const double*)base)[t];
        const int32_t* ints = (const int32_t*)(base + chunk * 8);
        tri[e] = ints[t];
        sq_i[e] = ints[chunk + t];
        sq_j[e] = ints[2 * chunk + t];
        gamma[e] = ints[3 * chunk + t];
    }
}
}  // namespace dcr

extern "C" int dcr_bfc_paper_unshard(const void* gathered, int world, int64_t chunk, int64_t n_edges, int32_t* out_tri,
                                     int32_t* out_sq_i, int32_t* out_sq_j, int32_t* out_gamma, double* out_bfc,
                                     void* stream) {
    if (n_edges <= 0) return 0;
    if (world <= 0 || chunk * world < n_edges) { set_error("dcr_bfc_paper_unshard: bad shard geometry"); return 1; }
    const int ctas = (int)std::min<int64_t>((n_edges + 255) / 256, (int64_t)sm_count() * 8);
    unshard_kernel<<<ctas, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)gathered, world, chunk, n_edges,
                                                          out_tri, out_sq_i, out_sq_j, out_gamma, out_bfc);
    DCR_LAUNCH_CHECK();
    return 0;
}
