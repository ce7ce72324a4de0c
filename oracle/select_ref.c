/*
 * oracle/select_ref.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never on the product path).
 *
 * C + OpenMP restatement of the native half of `implicit.cpu.topk.topk`
 * (pm-implicit 0.7.3, third-party, not vendored under /root/reference): after
 * the BLAS product `scores = query . items^T` (done by the caller with numpy,
 * as the upstream Cython does per query batch), every row is
 *   1. divided by `item_norms` when given          (call site rank_implicit.py:268),
 *   2. masked with -FLT_MAX at the stored entries of the CSR `filter_query_items`
 *      (rank_implicit.py:269; sentinel contract tests/models/rank/test_rank_implicit.py:51-71),
 *   3. reduced to its k best (score, id) pairs, sorted by score descending,
 * rows in parallel over `num_threads` OpenMP threads (rank_implicit.py:271).
 * Tie order is implementation-defined upstream; here (score desc, id asc).
 *
 * Build: see oracle/Makefile (gcc -O3 -fopenmp -shared -fPIC).
 */
#include <float.h>
#include <stdint.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    float s;
    int32_t id;
} pair_t;

/* "a ranks before b": higher score first, then smaller id */
static inline int before(pair_t a, pair_t b) { return a.s > b.s || (a.s == b.s && a.id < b.id); }

/* min-heap on rank order: root = worst kept element */
static inline void sift_down(pair_t* h, int n, int i) {
    for (;;) {
        int l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && before(h[w], h[l])) w = l;
        if (r < n && before(h[w], h[r])) w = r;
        if (w == i) return;
        pair_t t = h[i];
        h[i] = h[w];
        h[w] = t;
        i = w;
    }
}

static int cmp_rank(const void* pa, const void* pb) {
    pair_t a = *(const pair_t*)pa, b = *(const pair_t*)pb;
    if (before(a, b)) return -1;
    if (before(b, a)) return 1;
    return 0;
}

/*
 * scores      [n_rows, n_items] fp32, row-major, MODIFIED in place (norm division + mask)
 * item_norms  [n_items] or NULL
 * indptr      [n_rows + 1] int64 or NULL (no filter); indices int32, column ids >= n_items ignored
 * out_ids     [n_rows, k] int32, out_scores [n_rows, k] fp32
 * returns 0, or -1 on bad arguments
 */
int ref_topk_select(float* scores, int64_t n_rows, int64_t n_items, const float* item_norms, const int64_t* indptr,
                    const int32_t* indices, int32_t k, int32_t* out_ids, float* out_scores, int32_t num_threads) {
    if (k <= 0 || k > n_items || n_rows < 0) return -1;
#ifdef _OPENMP
    if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
#pragma omp parallel
    {
        pair_t* heap = (pair_t*)malloc(sizeof(pair_t) * (size_t)k);
#pragma omp for schedule(dynamic, 8)
        for (int64_t r = 0; r < n_rows; ++r) {
            float* row = scores + r * n_items;
            if (item_norms)
                for (int64_t j = 0; j < n_items; ++j) row[j] /= item_norms[j];
            if (indptr)
                for (int64_t p = indptr[r]; p < indptr[r + 1]; ++p)
                    if (indices[p] >= 0 && indices[p] < n_items) row[indices[p]] = -FLT_MAX;
            int n = 0;
            for (int64_t j = 0; j < n_items; ++j) {
                pair_t c = {row[j], (int32_t)j};
                if (n < k) {
                    heap[n++] = c;
                    if (n == k)
                        for (int i = k / 2 - 1; i >= 0; --i) sift_down(heap, k, i);
                } else if (before(c, heap[0])) {
                    heap[0] = c;
                    sift_down(heap, k, 0);
                }
            }
            qsort(heap, (size_t)n, sizeof(pair_t), cmp_rank);
            for (int i = 0; i < k; ++i) {
                out_ids[r * k + i] = heap[i].id;
                out_scores[r * k + i] = heap[i].s;
            }
        }
        free(heap);
    }
    return 0;
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
