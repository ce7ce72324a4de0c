// Exhaustive fp64 scoring kernel + warp-level list selection / re-scoring kernels.
//
// Replaces, for the rows it is given, the arithmetic of `implicit.cpu.topk.topk` as called at
// rectools/models/rank/rank_implicit.py:264-272 (score, /item_norms, CSR mask, per-row top-k), with the result
// definition of include/b200_rank.h (fp64-accumulated dot rounded once to fp32; order = score desc, id asc).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int LIST_LEN = 32;  // every partial list is one warp wide

// ---------------------------------------------------------------------------------------------------------------
// Exhaustive kernel: 32 subjects per block (4 per warp), object positions streamed 32 at a time (lane = object).
// ---------------------------------------------------------------------------------------------------------------
constexpr int EX_THREADS = 256;
constexpr int EX_ROWS = 32;
constexpr int EX_ROWS_PER_WARP = 4;
constexpr int EX_DK = 64;

struct ExactParams {
    const float* subjects;   // fp32 [*, d]
    const int64_t* row_map;  // nullable: logical row -> physical row of `subjects`
    const int32_t* rows;     // nullable: compact index -> logical row (re-rank subset)
    const int32_t* n_sel_dev;  // nullable: device-side count overriding n_sel (early exit for unused blocks)
    int64_t n_sel;           // number of compact indices
    const float* objects;    // fp32 [n_objects, d]
    const int32_t* pos2obj;  // nullable whitelist: position -> object id (sorted ascending)
    int64_t n_pos;
    int32_t d;
    const float* obj_norms;  // nullable (COSINE): fp32 norm per object id, zero already replaced by 1e-10
    const int64_t* indptr;   // nullable CSR filter, by logical row
    const int32_t* indices;
    int32_t id_off;          // CSR column ids are global: global id = local object id + id_off
    int32_t k_out;           // row stride of out_*
    int32_t k0;              // this pass selects entries [k0, k0 + kp)
    int32_t kp;              // <= 32
    const int32_t* out_ids;  // previous passes (bound), [n_rows, k_out]
    const float* out_scores;
    const int32_t* out_counts;
    float* part_scores;  // [n_splits][part_stride_rows][32]
    int32_t* part_ids;
    int64_t part_stride_rows;  // >= n_sel
};

__global__ void __launch_bounds__(EX_THREADS) exact_topk_kernel(const ExactParams p) {
    __shared__ float s_obj[32][EX_DK + 1];
    __shared__ float s_sub[EX_ROWS][EX_DK];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t n_sel = p.n_sel_dev ? (int64_t)*p.n_sel_dev : p.n_sel;
    const int64_t sel0 = (int64_t)blockIdx.x * EX_ROWS;
    if (sel0 >= n_sel) return;

    const int n_splits = gridDim.y, split = blockIdx.y;
    const int64_t tiles_total = (p.n_pos + 31) >> 5;
    const int64_t tiles_per = (tiles_total + n_splits - 1) / n_splits;
    const int64_t t0 = split * tiles_per;
    const int64_t t1 = min(t0 + tiles_per, tiles_total);

    // per-warp subjects
    int64_t lrow[EX_ROWS_PER_WARP], flo[EX_ROWS_PER_WARP], fhi[EX_ROWS_PER_WARP];
    bool active[EX_ROWS_PER_WARP];
    float thr[EX_ROWS_PER_WARP], ls[EX_ROWS_PER_WARP], bs[EX_ROWS_PER_WARP];
    int li[EX_ROWS_PER_WARP], bi[EX_ROWS_PER_WARP];
#pragma unroll
    for (int q = 0; q < EX_ROWS_PER_WARP; ++q) {
        const int64_t sel = sel0 + warp * EX_ROWS_PER_WARP + q;
        active[q] = sel < n_sel;
        lrow[q] = active[q] ? (p.rows ? (int64_t)p.rows[sel] : sel) : 0;
        thr[q] = -INFINITY;
        ls[q] = -INFINITY;
        li[q] = B200_PAD_ID;
        bs[q] = INFINITY;
        bi[q] = -1;
        flo[q] = fhi[q] = 0;
        if (active[q]) {
            if (p.indptr) {
                flo[q] = p.indptr[lrow[q]];
                fhi[q] = p.indptr[lrow[q] + 1];
            }
            if (p.k0 > 0) {
                if (p.out_counts[lrow[q]] < p.k0) {
                    active[q] = false;  // row exhausted by earlier passes
                } else {
                    bs[q] = p.out_scores[lrow[q] * p.k_out + p.k0 - 1];
                    bi[q] = p.out_ids[lrow[q] * p.k_out + p.k0 - 1];
                }
            }
        }
    }

    for (int64_t t = t0; t < t1; ++t) {
        double acc[EX_ROWS_PER_WARP];
#pragma unroll
        for (int q = 0; q < EX_ROWS_PER_WARP; ++q) acc[q] = 0.0;

        for (int dk0 = 0; dk0 < p.d; dk0 += EX_DK) {
            __syncthreads();
            for (int e = tid; e < 32 * EX_DK; e += EX_THREADS) {
                const int it = e >> 6, j = e & (EX_DK - 1);
                const int64_t pos = t * 32 + it;
                float v = 0.f;
                if (pos < p.n_pos && dk0 + j < p.d) {
                    const int64_t obj = p.pos2obj ? (int64_t)p.pos2obj[pos] : pos;
                    v = __ldg(p.objects + obj * p.d + dk0 + j);
                }
                s_obj[it][j] = v;
            }
            for (int e = tid; e < EX_ROWS * EX_DK; e += EX_THREADS) {
                const int r = e >> 6, j = e & (EX_DK - 1);
                const int64_t sel = sel0 + r;
                float v = 0.f;
                if (sel < n_sel && dk0 + j < p.d) {
                    const int64_t lr = p.rows ? (int64_t)p.rows[sel] : sel;
                    const int64_t pr = p.row_map ? p.row_map[lr] : lr;
                    v = __ldg(p.subjects + pr * p.d + dk0 + j);
                }
                s_sub[r][j] = v;
            }
            __syncthreads();
            const int jn = min(EX_DK, p.d - dk0);
            for (int j = 0; j < jn; ++j) {
                const double ov = (double)s_obj[lane][j];
#pragma unroll
                for (int q = 0; q < EX_ROWS_PER_WARP; ++q)
                    acc[q] = fma(ov, (double)s_sub[warp * EX_ROWS_PER_WARP + q][j], acc[q]);
            }
        }

        const int64_t pos = t * 32 + lane;
        const bool valid = pos < p.n_pos;
        const int obj = valid ? (p.pos2obj ? p.pos2obj[pos] : (int)pos) : B200_PAD_ID;
        const double inv_div = (valid && p.obj_norms) ? (double)__ldg(p.obj_norms + obj) : 1.0;
#pragma unroll
        for (int q = 0; q < EX_ROWS_PER_WARP; ++q) {
            if (!active[q]) continue;  // warp-uniform
            const float s = p.obj_norms ? (float)(acc[q] / inv_div) : (float)acc[q];
            bool c = valid && s > thr[q] && (s < bs[q] || (s == bs[q] && obj > bi[q]));
            if (c && p.indptr) c = !csr_contains(p.indices, flo[q], fhi[q], obj + p.id_off);
            unsigned m = __ballot_sync(B200_FULL_MASK, c);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cs = __shfl_sync(B200_FULL_MASK, s, src);
                const int ci = __shfl_sync(B200_FULL_MASK, obj, src);
                if (!(cs > thr[q])) continue;  // threshold rose while draining this ballot
                const int ins = __popc(__ballot_sync(B200_FULL_MASK, ls[q] >= cs));
                const float us = __shfl_up_sync(B200_FULL_MASK, ls[q], 1);
                const int ui = __shfl_up_sync(B200_FULL_MASK, li[q], 1);
                if (lane == ins) {
                    ls[q] = cs;
                    li[q] = ci;
                } else if (lane > ins) {
                    ls[q] = us;
                    li[q] = ui;
                }
                thr[q] = __shfl_sync(B200_FULL_MASK, ls[q], p.kp - 1);
            }
        }
    }

#pragma unroll
    for (int q = 0; q < EX_ROWS_PER_WARP; ++q) {
        const int64_t sel = sel0 + warp * EX_ROWS_PER_WARP + q;
        if (sel >= n_sel) continue;
        const int64_t o = ((int64_t)split * p.part_stride_rows + sel) * LIST_LEN + lane;
        const bool keep = active[q] && lane < p.kp;
        p.part_scores[o] = keep ? ls[q] : -INFINITY;
        p.part_ids[o] = keep ? li[q] : B200_PAD_ID;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Selection kernel: one warp per row merges `n_lists` candidate lists into the next `kp` output entries.
// RESCORE = true: candidates carry approximate (tensor-core) scores; every candidate is re-scored in fp64 from the
// fp32 master copies, and the row is certified or queued for the exhaustive kernel.
// ---------------------------------------------------------------------------------------------------------------
struct SelectParams {
    const float* in_scores;    // [n_lists][n_sel][L]
    const int32_t* in_ids;     // pad entries: id == B200_PAD_ID (or beyond in_counts)
    const int32_t* in_counts;  // nullable [n_lists][n_sel]
    int32_t n_lists;
    int32_t L;
    int64_t n_sel;
    int64_t list_stride_rows;  // rows between consecutive lists (>= n_sel)
    const int32_t* rows;       // nullable: compact index -> logical row
    const int32_t* n_sel_dev;  // nullable device-side n_sel
    int32_t k_out, k0, kp;
    int32_t* out_ids;
    float* out_scores;
    int32_t* out_counts;
    // re-scoring inputs
    const float* subjects;
    const int64_t* row_map;
    const float* objects;
    const float* obj_norms;
    int32_t d;
    // certificate (RESCORE only)
    int32_t k_cand;           // a list with k_cand valid entries may have discarded objects
    const int32_t* row_exp;   // per-row power-of-two exponent applied to the subject before rounding
    int32_t obj_exp;          // exponent applied to the objects
    float eps_rel;            // bound on |approx - exact| / (|u|_2 * max_i |i|_2), see DESIGN.md
    float max_obj_norm;       // max_i |i|_2 (1 for pre-normalised COSINE objects)
    int32_t* fb_count;        // device counter of rows that failed the certificate
    int32_t* fb_rows;         // their logical rows
};

constexpr int SEL_WARPS = 8;

template <bool RESCORE>
__global__ void __launch_bounds__(SEL_WARPS * 32) select_kernel(const SelectParams p) {
    extern __shared__ float s_sub[];  // RESCORE: [SEL_WARPS][d] subject rows
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_sel = p.n_sel_dev ? (int64_t)*p.n_sel_dev : p.n_sel;
    const int64_t sel = (int64_t)blockIdx.x * SEL_WARPS + warp;
    if (sel >= n_sel) return;
    const int64_t lrow = p.rows ? (int64_t)p.rows[sel] : sel;

    float bs = INFINITY;
    int bi = -1;
    if (p.k0 > 0) {
        if (p.out_counts[lrow] < p.k0) return;
        bs = p.out_scores[lrow * p.k_out + p.k0 - 1];
        bi = p.out_ids[lrow * p.k_out + p.k0 - 1];
    }

    float* sub = nullptr;
    double unorm2 = 0.0;
    if (RESCORE) {
        sub = s_sub + (size_t)warp * p.d;
        const int64_t pr = p.row_map ? p.row_map[lrow] : lrow;
        for (int j = lane; j < p.d; j += 32) {
            const float v = __ldg(p.subjects + pr * p.d + j);
            sub[j] = v;
            unorm2 = fma((double)v, (double)v, unorm2);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) unorm2 += __shfl_xor_sync(B200_FULL_MASK, unorm2, o);
        __syncwarp();
    }

    float run_s = -INFINITY;
    int run_i = B200_PAD_ID;
    int n_valid = 0;
    float thr_approx = -INFINITY;  // best "K'-th approximate score" over the full lists
    bool any_full = false;
    const int total = p.n_lists * p.L;
    for (int base = 0; base < total; base += 32) {
        const int c = base + lane;
        const int list = c / p.L, e = c - list * p.L;
        bool valid = c < total;
        const int64_t o = valid ? ((int64_t)list * p.list_stride_rows + sel) * p.L + e : 0;
        int id = B200_PAD_ID;
        float s = -INFINITY;
        int cnt = p.L;
        if (valid) {
            if (p.in_counts) cnt = p.in_counts[(int64_t)list * p.list_stride_rows + sel];
            valid = e < cnt;
            if (valid) id = p.in_ids[o];
            valid = valid && id != B200_PAD_ID && id >= 0;
        }
        if (RESCORE) {
            // certificate bookkeeping on the approximate scores
            // (lists are unsorted and L == 32: one list per 32-lane chunk; a full list's smallest approximate score
            // bounds everything that list ever discarded)
            float a = valid ? p.in_scores[o] : INFINITY;
#pragma unroll
            for (int sh = 16; sh > 0; sh >>= 1) a = fminf(a, __shfl_xor_sync(B200_FULL_MASK, a, sh));
            if (cnt >= p.k_cand) {
                any_full = true;
                thr_approx = fmaxf(thr_approx, a);
            }
            double acc = 0.0;
            const float* orow = p.objects + (int64_t)(valid ? id : 0) * p.d;
            if ((p.d & 3) == 0) {
                const float4* o4 = reinterpret_cast<const float4*>(orow);
                const float4* s4 = reinterpret_cast<const float4*>(sub);
                for (int j = 0; j < (p.d >> 2); ++j) {
                    const float4 ov = __ldg(o4 + j);
                    const float4 sv = s4[j];
                    acc = fma((double)ov.x, (double)sv.x, acc);
                    acc = fma((double)ov.y, (double)sv.y, acc);
                    acc = fma((double)ov.z, (double)sv.z, acc);
                    acc = fma((double)ov.w, (double)sv.w, acc);
                }
            } else {
                for (int j = 0; j < p.d; ++j) acc = fma((double)__ldg(orow + j), (double)sub[j], acc);
            }
            s = p.obj_norms ? (float)(acc / (double)__ldg(p.obj_norms + (valid ? id : 0))) : (float)acc;
        } else if (valid) {
            s = p.in_scores[o];
        }
        valid = valid && (s < bs || (s == bs && id > bi));
        if (!valid) {
            s = -INFINITY;
            id = B200_PAD_ID;
        }
        n_valid += __popc(__ballot_sync(B200_FULL_MASK, valid));
        warp_sort32(s, id, lane);
        warp_merge_top32(run_s, run_i, s, id, lane);
    }

    const int n_out = min(n_valid, p.kp);
    if (lane < p.kp) {
        const bool w = lane < n_out;
        p.out_ids[lrow * p.k_out + p.k0 + lane] = w ? run_i : -1;
        p.out_scores[lrow * p.k_out + p.k0 + lane] = w ? run_s : -FLT_MAX;
    }
    if (lane == 0) p.out_counts[lrow] = p.k0 + n_out;

    if (RESCORE) {
        any_full = __any_sync(B200_FULL_MASK, any_full);
        if (any_full) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) thr_approx = fmaxf(thr_approx, __shfl_xor_sync(B200_FULL_MASK, thr_approx, o));
            const float e_k = __shfl_sync(B200_FULL_MASK, run_s, p.kp - 1);
            // approximate scores are in units scaled by 2^(row_exp + obj_exp)
            const int ex = (p.row_exp ? p.row_exp[sel] : 0) + p.obj_exp;  // row_exp is indexed by batch row
            const double thr = ldexp((double)thr_approx, -ex);
            const double eps = (double)p.eps_rel * sqrt(unorm2) * (double)p.max_obj_norm;
            // one fp32 ulp of slack: a discarded object whose exact score rounds up to e_k could tie with a smaller id
            const bool ok = n_valid >= p.kp && (double)e_k > thr + eps + 1.2e-7 * fabs((double)e_k);
            if (!ok && lane == 0) {
                const int slot = atomicAdd(p.fb_count, 1);
                p.fb_rows[slot] = (int32_t)lrow;
            }
        }
    }
}

// Multi-pass ranking (k > 24 on the tensor-core path): the ids a row has received so far, sorted ascending, become the
// row's exclusion list for the next pass (one warp per row, rank counting; unfilled slots sort last as B200_PAD_ID).
__global__ void build_exclusion_kernel(const int32_t* __restrict__ out_ids, int64_t n_rows, int32_t k_out, int32_t k0,
                                       int32_t id_off, int32_t* __restrict__ excl) {
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const int32_t* src = out_ids + row * k_out;
    for (int i = lane; i < k0; i += 32) {
        const int32_t raw = src[i];
        const int32_t key = raw < 0 ? B200_PAD_ID : raw + id_off;  // the kernels compare GLOBAL ids
        int rank = 0;
        for (int j = 0; j < k0; ++j) {
            const int32_t rj = src[j];
            const int32_t kj = rj < 0 ? B200_PAD_ID : rj + id_off;
            rank += (kj < key || (kj == key && j < i)) ? 1 : 0;
        }
        excl[row * k_out + rank] = key;
    }
}

// Local object ids -> global ids of an item-sharded catalogue (unfilled slots stay -1).
__global__ void add_offset_kernel(int32_t* ids, int64_t n, int32_t off) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && ids[i] >= 0) ids[i] += off;
}

// Initialise the output arrays: ids = -1, scores = -FLT_MAX, counts = 0.
__global__ void init_outputs_kernel(int32_t* ids, float* scores, int32_t* counts, int64_t n_rows, int32_t k_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows * k_out) {
        ids[i] = -1;
        scores[i] = -FLT_MAX;
    }
    if (i < n_rows) counts[i] = 0;
}

}  // namespace b200
