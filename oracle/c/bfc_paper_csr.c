/*
 * Oracle (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py): plain-C restatement of the paper-flavour Balanced
 * Forman curvature, used (a) as a second checker at sizes where the Python-set oracle is too slow and (b) as the
 * timed CPU baseline of bench.py (`cpu_baseline`, `--impl reference`).  Never linked into the product library.
 *
 * Follows /root/reference/curvature/bfc_naive.py:
 *   bfc_edge :7-40   one call of edge_fields() below, same set definitions, value evaluated left to right
 *   bfc      :43-52  the loop over edges in oracle_bfc_paper()
 * Sets are realised with a per-thread marker array over node ids (bit 0: in N(v1), bit 1: in N(v2)).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

static void edge_fields(const int32_t* rowptr, const int32_t* col, uint8_t* mark, int v1, int v2, int32_t* out4,
                        double* value) {
    const int b1 = rowptr[v1], e1 = rowptr[v1 + 1], b2 = rowptr[v2], e2 = rowptr[v2 + 1];
    const int deg1 = e1 - b1, deg2 = e2 - b2;                      /* :15-16 */
    const int deg_min = deg1 < deg2 ? deg1 : deg2;                 /* :17    */
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (deg_min == 1) { *value = 0.0; return; }                    /* :18-19 */
    const int deg_max = deg1 > deg2 ? deg1 : deg2;                 /* :20    */
    for (int p = b1; p < e1; ++p) mark[col[p]] |= 1;               /* S1_1   :22 */
    for (int p = b2; p < e2; ++p) mark[col[p]] |= 2;               /* S1_2   :23 */
    int tri = 0;
    for (int p = b1; p < e1; ++p) tri += mark[col[p]] == 3;        /* :25    */
    /* squares_1 = {k in S1_1 - S1_2, k != v2 : (N(k) & S1_2) - (S1_1 | {v1}) non-empty}     :26-27
     * gamma term for k: |N(k) & (S1_2 - S1_1)| - 1  (the -1 is v1 itself)                    :36     */
    int sq1 = 0, sq2 = 0, g = 0;
    for (int p = b1; p < e1; ++p) {
        const int k = col[p];
        if (mark[k] != 1 || k == v2) continue;
        int c = 0;
        for (int q = rowptr[k]; q < rowptr[k + 1]; ++q) { const int w = col[q]; c += (mark[w] == 2 && w != v1); }
        if (c > 0) { ++sq1; if (c > g) g = c; }
    }
    for (int p = b2; p < e2; ++p) {                                /* squares_2 mirrored   :28-29, :37 */
        const int k = col[p];
        if (mark[k] != 2 || k == v1) continue;
        int c = 0;
        for (int q = rowptr[k]; q < rowptr[k + 1]; ++q) { const int w = col[q]; c += (mark[w] == 1 && w != v2); }
        if (c > 0) { ++sq2; if (c > g) g = c; }
    }
    for (int p = b1; p < e1; ++p) mark[col[p]] = 0;
    for (int p = b2; p < e2; ++p) mark[col[p]] = 0;
    double t = 2.0 / deg1 + 2.0 / deg2 - 2.0 + (double)(2 * (long long)tri) / deg_max + (double)tri / deg_min;
    int gamma = 0;
    if (sq1 != 0 && sq2 != 0) {                                    /* :30-32 vs :34-40 */
        gamma = g;
        t = t + 1.0 / gamma / deg_max * (double)(sq1 + sq2);
    }
    out4[0] = tri; out4[1] = sq1; out4[2] = sq2; out4[3] = gamma;
    *value = t;
}

/* Returns 0 on success.  Outputs are indexed by position in the (esrc, edst) list.  Edges are handed out to
 * `nthreads` pthreads in blocks of 16 through an atomic counter (libgomp is not in the image). */
typedef struct {
    const int32_t *rowptr, *col, *esrc, *edst;
    int n;
    int64_t m;
    int32_t *tri, *sq1, *sq2, *gamma;
    double* value;
    int64_t* next;
    int* failed;
} job_t;

static void* worker(void* arg) {
    job_t* j = (job_t*)arg;
    uint8_t* mark = (uint8_t*)calloc((size_t)j->n + 1, 1);
    if (!mark) { __atomic_store_n(j->failed, 1, __ATOMIC_RELAXED); return NULL; }
    for (;;) {
        const int64_t lo = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
        if (lo >= j->m) break;
        const int64_t hi = lo + 16 < j->m ? lo + 16 : j->m;
        for (int64_t e = lo; e < hi; ++e) {
            int32_t f[4];
            double v;
            edge_fields(j->rowptr, j->col, mark, j->esrc[e], j->edst[e], f, &v);
            j->tri[e] = f[0]; j->sq1[e] = f[1]; j->sq2[e] = f[2]; j->gamma[e] = f[3]; j->value[e] = v;
        }
    }
    free(mark);
    return NULL;
}

int oracle_bfc_paper(const int32_t* rowptr, const int32_t* col, int n, const int32_t* esrc, const int32_t* edst,
                     int64_t m, int32_t* tri, int32_t* sq1, int32_t* sq2, int32_t* gamma, double* value,
                     int nthreads) {
    int failed = 0;
    int64_t next = 0;
    job_t job = {rowptr, col, esrc, edst, n, m, tri, sq1, sq2, gamma, value, &next, &failed};
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
    if (!th) return 1;
    int started = 0;
    for (int t = 1; t < nthreads; ++t)
        if (pthread_create(&th[started], NULL, worker, &job) == 0) ++started;
    worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    return failed;
}
