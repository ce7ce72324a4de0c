/*
 * b200_rank.h -- C ABI of the B200-native score + top-K engine (libb200rank.so).
 *
 * This is the drop-in boundary for RecTools' vector-ranking hot path.  Paths below are relative to the reference
 * checkout (RecTools 0.17.0).  Plain pointers and sizes only; no Python / torch types cross this boundary.
 *
 * What each entry point replaces in the reference:
 *
 *   b200_rank_create / b200_rank_destroy
 *       `ImplicitRanker.__init__` object-factor handling (rectools/models/rank/rank_implicit.py:58-81) and the per-call
 *       upload of the whole item matrix by `implicit.gpu.Matrix` (rank_implicit.py:156, rectools/models/utils.py:136);
 *       `TorchRanker.__init__` / `item_embs.to(device)` (rank_torch.py:59-75, :135).  The engine keeps the object
 *       factors resident in HBM (fp32 master copy + fp16/bf16 tensor-core copy + fp32 row norms for COSINE,
 *       rank_implicit.py:98-105, :238-240).
 *   b200_rank_set_subjects
 *       `self.subjects_factors = subjects_factors.astype(np.float32)` (rank_implicit.py:70) -- resident subject
 *       factors so that `rank(subject_ids=...)` gathers rows on the device (rank_implicit.py:236).
 *   b200_rank_topk
 *       the third-party call `implicit.cpu.topk.topk(items, query, k, item_norms, filter_query_items, ...)`
 *       (rank_implicit.py:264-272) and `implicit.gpu.KnnQuery().topk(...)` (rank_implicit.py:175-182), fused with the
 *       whitelist gather / CSR column restriction (rank_implicit.py:219-226), the whitelist id remap (:274-275) and
 *       the trailing-sentinel strip of `_process_implicit_scores` (:107-118): filtered items are never returned and
 *       `out_counts[r]` gives the number of valid leading entries of row r.  Also replaces the batched
 *       `scores = user_embs @ item_embs.T; masked_fill; torch.topk` loop of `TorchRanker.rank` (rank_torch.py:122-155).
 *   b200_rank_merge
 *       no reference counterpart (the reference is single-device); merges per-shard top-K lists after the NCCL
 *       all-gather of an item-sharded catalogue (BASELINE.json north_star).
 *
 * Result definition (the oracle, oracle/topk_oracle.py `accum="f64"`): score(u, i) = fp32( sum_j fp64(u_j) * fp64(i_j) )
 * for DOT; for COSINE fp32( dot64 / fp64(norm_i) ) with norm_i = fp32(sqrt(sum_j fp64(i_j)^2)), zero -> 1e-10
 * (the division by the subject norm is left to the caller exactly as in rank_implicit.py:132-134).  Rows are ordered
 * by (score descending, object id ascending).  The tensor-core path only proposes candidates; every returned score is
 * re-computed as defined above and every row is either certified (no discarded object can enter the top-k) or
 * re-ranked by the exhaustive fp64 kernel, so results do not depend on the path taken.
 *
 * Error convention: every function returns 0 on success or a negative B200_E_* code; a human-readable message for the
 * last failure on the calling thread is available from b200_rank_last_error().  There is NO CPU fallback: if no
 * sm_100 device is present b200_rank_create fails with B200_E_CUDA.
 *
 * Threading: calls on one engine are serialised by an internal mutex; distinct engines are independent.
 */
#ifndef B200_RANK_H
#define B200_RANK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_RANK_ABI_VERSION 3

/* error codes */
#define B200_OK 0
#define B200_E_INVALID (-1) /* contract violation (bad shape / pointer / k) */
#define B200_E_CUDA (-2)    /* CUDA runtime / driver failure, or no sm_100 device */
#define B200_E_NOMEM (-3)
#define B200_E_UNSUPPORTED (-4)

/* distance (rectools/models/rank/rank.py:25-30); EUCLIDEAN is served by the caller through the dot-augmentation
 * trick exactly as the reference does (rank_implicit.py:242-246, :136-140) */
#define B200_DIST_DOT 0
#define B200_DIST_COSINE 1

/* tensor-core candidate pass */
#define B200_TC_AUTO 0 /* fp16 (power-of-two scaled) when the factors fit its range, else bf16 */
#define B200_TC_FP16 1
#define B200_TC_BF16 2
#define B200_TC_OFF 3 /* exhaustive fp64 kernel only */

/* create flags */
#define B200_F_OBJECTS_ON_DEVICE 1 /* `objects` is a device pointer on `device` */

/* query flags */
#define B200_Q_INPUTS_ON_DEVICE 1  /* subjects / subject_ids / csr_* / whitelist are device pointers */
#define B200_Q_OUTPUTS_ON_DEVICE 2 /* out_* are device pointers */
#define B200_Q_FORCE_EXACT 4       /* skip the tensor-core pass */
#define B200_Q_FORCE_TC 8          /* fail with B200_E_UNSUPPORTED instead of silently using the exhaustive kernel */
#define B200_Q_SHARED_THRESHOLDS 16 /* item-sharded multi-GPU pass: prune with the maximum of all ranks' thresholds (peer memory
                                     * set up by b200_rank_peer_*), no local verdict: `out_bounds` must be given and the global
                                     * top-k is certified by b200_rank_merge_certified.  Needs k <= 24. */

/* element types of factor matrices handed over as device pointers (b200_rank_create_ex / query.subject_dtype) */
#define B200_DT_F32 0
#define B200_DT_F16 1
#define B200_DT_BF16 2

typedef struct b200_rank_engine b200_rank_engine;

typedef struct b200_rank_query {
    /* subjects to rank.  Either `subjects` ([n_rows, d] fp32, row-major) or, when NULL, `subject_ids` indexing the
     * matrix given to b200_rank_set_subjects.  If both are given, row r is subjects[subject_ids[r]]. */
    const float* subjects;
    const int64_t* subject_ids;
    int64_t n_rows;
    int64_t n_subjects_total; /* rows in `subjects` when subject_ids is given with an explicit matrix, else 0 */
    /* filter_pairs_csr (rank.py:38): structure only, row r = stored column ids (sorted ascending within a row,
     * int32, global object ids; ids >= n_objects are ignored).  NULL indptr = no filter. */
    const int64_t* csr_indptr; /* [n_rows + 1] */
    const int32_t* csr_indices;
    /* sorted_object_whitelist (rank.py:39): sorted unique object ids, or NULL */
    const int32_t* whitelist;
    int64_t n_whitelist;
    int32_t k;     /* requested k; the engine uses real_k = min(k, n candidates) (rank_implicit.py:248) */
    int32_t flags; /* B200_Q_* */
    /* outputs, [n_rows, k_out] row-major with k_out = min(k, n_whitelist or n_objects); unfilled slots hold id = -1,
     * score = -FLT_MAX */
    int32_t* out_ids;
    float* out_scores;
    int32_t* out_counts; /* [n_rows] */
    void* stream;        /* cudaStream_t the device buffers are produced / consumed on; the call is ordered after the work
                          * queued on it and it waits for the results.  NULL = the (legacy) default stream.  Ignored when
                          * every buffer is a host buffer. */
    /* ---- ABI 3 */
    float* out_bounds;   /* B200_Q_SHARED_THRESHOLDS: [n_rows] upper bound on the exact score of every object of this shard that
                          * is NOT among the row's returned candidates (-inf: nothing was discarded); same memory space as out_* */
    uint32_t peer_epoch; /* B200_Q_SHARED_THRESHOLDS: tag of this call, >= 1, the same on every rank, different from the
                          * previous call's */
    int32_t subject_dtype; /* B200_DT_* of `subjects` (device pointers only; host matrices are fp32) */
    /* sparse subjects (EASEModel: subjects_factors is the user x item CSR, rectools/models/ease.py:134-161, DOT only):
     * row r of the batch = sub_indices / sub_data [sub_indptr[r], sub_indptr[r+1]); columns index the d factor columns.
     * Given instead of `subjects` / `subject_ids`. */
    const int64_t* sub_indptr; /* [n_rows + 1] or NULL */
    const int32_t* sub_indices;
    const float* sub_data;
    int64_t reserved[2];
} b200_rank_query;

typedef struct b200_rank_stats {
    int32_t path;            /* 0 = exhaustive fp64 kernel, 1 = tensor-core candidates + fp64 re-score, 2 = sparse subjects (SpMM
                              * scores + streaming selection), 3 = k > 128: exhaustive scores materialised once + selection passes */
    int32_t tc_dtype;        /* B200_TC_FP16 / B200_TC_BF16 when path == 1 */
    int32_t k_out;           /* columns of the output arrays */
    int32_t k_cand;          /* candidates kept per row and item split by the tensor-core pass */
    int32_t n_splits;        /* item splits of the main kernel */
    int32_t n_launches;      /* kernels launched by this call */
    int64_t n_fallback_rows; /* rows whose certificate failed after the first tensor-core pass (re-ranked with wider lists) */
    int64_t n_exact_rows;    /* rows that still failed and were ranked by the exhaustive fp64 kernel */
    float ms_main;           /* CUDA-event time of the dominant kernel (tensor-core pass or exhaustive kernel), summed over chunks */
    float ms_total;          /* CUDA-event time of the whole call on the engine stream (copies included) */
    float ms_h2d;            /* exposed host->device staging inside ms_total (first chunk; later chunks overlap with compute) */
    float ms_d2h;            /* exposed device->host copy (last chunk) */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int32_t n_chunks;        /* row chunks of the copy / compute pipeline (1: call not chunked) */
    int32_t n_tc_launches;   /* launches of the fused tensor-core kernel summed in ms_main (main pass, second chance, re-rank passes) */
    int32_t epi_warps;       /* epilogue warps per CTA of the fused kernel (8 or 16) */
    int32_t wide;            /* 1: single-pass wide mode (24 < k <= 128) */
    float ms_select;         /* CUDA-event time of the fp64 re-score / selection kernels */
    int32_t reserved;
} b200_rank_stats;

typedef struct b200_rank_info {
    int32_t abi_version;
    int32_t device;
    int32_t sm_count;
    int32_t cc_major;
    int32_t cc_minor;
    int32_t tc_dtype; /* resolved tensor-core dtype of the engine (B200_TC_*) */
    int64_t n_objects;
    int32_t d;
    int32_t d_pad;
    int64_t hbm_bytes; /* device memory held by the engine */
    char device_name[128];
} b200_rank_info;

int b200_rank_create(b200_rank_engine** out, const float* objects, int64_t n_objects, int32_t d, int32_t distance,
                     int32_t device, int32_t tc_mode, int32_t flags);
/* The same with an explicit element type: fp16 / bf16 object factors (transformer id-embedding scorers keep `item_embs`
 * in the model dtype, rectools/models/nn/transformers/lightning.py:391-398).  16-bit matrices must be device pointers
 * (B200_F_OBJECTS_ON_DEVICE); they are widened once into the engine's fp32 master copy, which is exact. */
int b200_rank_create_ex(b200_rank_engine** out, const void* objects, int32_t dtype, int64_t n_objects, int32_t d,
                        int32_t distance, int32_t device, int32_t tc_mode, int32_t flags);
int b200_rank_destroy(b200_rank_engine* engine);
int b200_rank_set_subjects(b200_rank_engine* engine, const float* subjects, int64_t n_subjects, int32_t on_device);
/* Item-sharded catalogues: the engine holds objects [offset, offset + n_objects) of a larger catalogue.  CSR column ids
 * and returned ids are GLOBAL (local + offset); whitelist entries stay LOCAL positions into this shard. */
int b200_rank_set_id_offset(b200_rank_engine* engine, int64_t offset);
int b200_rank_topk(b200_rank_engine* engine, const b200_rank_query* query, b200_rank_stats* stats /* nullable */);
int b200_rank_get_info(b200_rank_engine* engine, b200_rank_info* info);

/* Merge `n_lists` per-shard results (device pointers, each [n_rows, k] / [n_rows], list l at base + l * stride) into
 * the global top-k ordered by (score desc, id asc).  Runs on `stream` of `device`. */
int b200_rank_merge(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids,
                    const float* scores, const int32_t* counts, int32_t* out_ids, float* out_scores,
                    int32_t* out_counts);

/* The same over per-shard results that carry certificate bounds (B200_Q_SHARED_THRESHOLDS passes): list l's arrays start
 * `list_stride` ELEMENTS (4-byte units) after list l-1's (one packed buffer per rank after an all-gather; 0 = dense arrays
 * as in b200_rank_merge).  Rows whose k-th merged score does not exceed every shard's bound are appended to `fail_rows`
 * (device, [n_rows]) and counted in `fail_count` (device int32, zeroed by the caller): they must be re-ranked without
 * threshold sharing. */
int b200_rank_merge_certified(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids,
                              const float* scores, const int32_t* counts, const float* bounds, int64_t list_stride,
                              int32_t* out_ids, float* out_scores, int32_t* out_counts, int32_t* fail_rows,
                              int32_t* fail_count);

/* Threshold sharing between the ranks of an item-sharded catalogue (one process per GPU, NVLink peer memory):
 *   export: allocate this engine's published-threshold array for calls of up to `max_rows` subject rows and return its
 *           64-byte CUDA IPC handle;
 *   import: open the arrays of all `n_ranks` ranks (`handles` = n_ranks x 64 bytes, in rank order; entry `self` is this
 *           engine's own and is skipped).  At most 9 ranks. */
int b200_rank_peer_export(b200_rank_engine* engine, int64_t max_rows, void* handle_out);
int b200_rank_peer_import(b200_rank_engine* engine, int32_t n_ranks, int32_t self, const void* handles);

const char* b200_rank_last_error(void);
int b200_rank_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* B200_RANK_H */
