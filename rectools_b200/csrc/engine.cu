// libb200rank.so -- host side of the B200 score + top-K engine behind the C ABI of include/b200_rank.h.
//
// Reference seams (RecTools 0.17.0): `ImplicitRanker.rank` (rectools/models/rank/rank_implicit.py:187-280),
// `ImplicitRanker._rank_on_gpu` (:148-185) and `TorchRanker.rank` (rectools/models/rank/rank_torch.py:77-177).
// The engine keeps the object factors resident (the reference re-uploads them per call, rank_implicit.py:156),
// stages one call's subjects / CSR filter / whitelist, runs the tensor-core candidate pass + fp64 re-score (or the
// exhaustive fp64 kernel) and returns padded [n_rows, k] arrays plus per-row counts.
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200_rank.h"
#include "common.cuh"
#include "prep.cuh"
#include "select.cuh"
#include "sparse.cuh"
#include "fused_topk.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

struct CudaError {
    cudaError_t e;
    const char* what;
    int line;
};

#define CK(call)                                               \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) throw CudaError{e__, #call, __LINE__}; \
    } while (0)

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    void ensure(size_t bytes) {
        if (bytes <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const {
        return reinterpret_cast<T*>(p);
    }
};

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_tiled() {
    static PFN_encodeTiled fn = nullptr;
    if (fn) return fn;
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !sym) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(sym);
    return fn;
}

// Row-major [rows, d_pad] 16-bit matrix, boxes of [128 rows x 64 cols] (128-byte rows, SWIZZLE_128B).
bool make_tensor_map(CUtensorMap* tm, const void* base, int64_t rows, int d_pad, bool is_bf16) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)b200::tc::KBLK, 128u};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}

}  // namespace

struct b200_rank_engine {
    std::mutex mu;
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    char dev_name[128] = {0};
    int distance = B200_DIST_DOT;
    int tc_dtype = B200_TC_FP16;  // resolved; B200_TC_OFF when the tensor-core path is unavailable
    int64_t n_obj = 0;
    int d = 0, d_pad = 0;
    int64_t n_obj_pad = 0;
    cudaStream_t st = nullptr;
    cudaStream_t cs = nullptr;          // copy stream of the chunk pipeline
    cudaEvent_t ev[8] = {nullptr};
    cudaEvent_t evp[3] = {nullptr};     // chunk pipeline: inputs of chunk c / c+1 staged, stream hand-over
    std::vector<cudaEvent_t> evt;       // timing pairs of the fused-kernel / selection launches of a call

    // resident object data
    DevBuf obj32;     // [n_obj, d] fp32 master copy
    DevBuf obj16;     // [n_obj_pad, d_pad] fp16 / bf16, pre-scaled (and pre-normalised for COSINE)
    DevBuf obj_norms; // [n_obj] fp32 (COSINE)
    DevBuf objT;      // [d, n_obj] fp32 transposed master copy (sparse subjects only, built on first use)
    int obj_exp = 0;
    int64_t id_offset = 0;
    float max_obj_norm = 0.f;
    const float* obj32_ptr = nullptr;

    // resident subjects (optional)
    DevBuf sub32_res;
    int64_t n_sub_res = 0;
    const float* sub32_res_ptr = nullptr;

    // threshold sharing with the other ranks of an item-sharded catalogue
    DevBuf peer_pub;
    int64_t peer_rows = 0;
    int n_peers = 0;
    void* peer_in[b200::tc::MAX_PEERS] = {nullptr};

    // per-call staging / workspace
    DevBuf sub32, sub16, row_exp, rowmap, indptr, indices, wl, obj16_wl;
    DevBuf sp_indptr, sp_indices, sp_data, sp_scores;
    DevBuf out_ids, out_scores, out_counts, out_bounds;
    DevBuf cand_scores, cand_ids, cand_counts, cand_thr;
    DevBuf part_scores, part_ids;
    DevBuf fb_rows, scratch, excl, carousel, patch;
    int32_t* h_pinned = nullptr;  // small pinned scratch (counters)
    std::vector<char> h_patch;    // host copy of re-ranked rows (host-output calls)

    std::vector<DevBuf*> all_bufs() {
        return {&obj32, &obj16, &obj_norms, &objT, &sub32_res, &peer_pub, &sub32, &sub16, &row_exp, &rowmap, &indptr, &indices, &wl,
                &obj16_wl, &sp_indptr, &sp_indices, &sp_data, &sp_scores, &out_ids, &out_scores, &out_counts, &out_bounds, &cand_scores,
                &cand_ids, &cand_counts, &cand_thr, &part_scores, &part_ids, &fb_rows, &scratch, &excl, &carousel, &patch};
    }
    size_t hbm_bytes() {
        size_t t = 0;
        for (auto* b : all_bufs()) t += b->cap;
        return t;
    }
    void free_all() {
        for (int i = 0; i < b200::tc::MAX_PEERS; ++i)
            if (peer_in[i]) cudaIpcCloseMemHandle(peer_in[i]);
        for (auto* b : all_bufs()) b->release();
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : evp)
            if (e) cudaEventDestroy(e);
        for (auto& e : evt)
            if (e) cudaEventDestroy(e);
        evt.clear();
        if (st) cudaStreamDestroy(st);
        if (cs) cudaStreamDestroy(cs);
        st = cs = nullptr;
    }
};

namespace {

using namespace b200;

int grid_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// ---- resident objects -------------------------------------------------------------------------------------
void prepare_objects(b200_rank_engine* E, int tc_mode) {
    const int64_t n = E->n_obj;
    const int d = E->d;
    E->scratch.ensure(64);
    unsigned* g = E->scratch.as<unsigned>();
    CK(cudaMemsetAsync(g, 0, 8, E->st));
    const bool cosine = E->distance == B200_DIST_COSINE;
    if (cosine) E->obj_norms.ensure(sizeof(float) * std::max<int64_t>(n, 1));
    if (n > 0)
        row_stats_kernel<<<grid_for(n * 32, 256), 256, 0, E->st>>>(E->obj32_ptr, n, d, cosine ? 1 : 0,
                                                                   cosine ? E->obj_norms.as<float>() : nullptr, g, g + 1);
    CK(cudaGetLastError());
    unsigned h[2];
    CK(cudaMemcpyAsync(h, g, 8, cudaMemcpyDeviceToHost, E->st));
    CK(cudaStreamSynchronize(E->st));
    float absmax, maxnorm;
    memcpy(&absmax, &h[0], 4);
    memcpy(&maxnorm, &h[1], 4);
    E->max_obj_norm = maxnorm;

    if (tc_mode == B200_TC_OFF || E->cc_major != 10 || E->sm_count % 2 != 0) {
        E->tc_dtype = B200_TC_OFF;
        return;
    }
    E->tc_dtype = (tc_mode == B200_TC_BF16) ? B200_TC_BF16 : B200_TC_FP16;
    E->obj_exp = (E->tc_dtype == B200_TC_FP16) ? fp16_scale_exp(absmax) : 0;
    E->n_obj_pad = round_up(std::max<int64_t>(n, 1), tc::HALF_N);
    E->obj16.ensure((size_t)E->n_obj_pad * E->d_pad * 2);
    const float* norms = cosine ? E->obj_norms.as<float>() : nullptr;
    const int grid = grid_for(E->n_obj_pad * 32, 256);
    if (E->tc_dtype == B200_TC_FP16)
        convert_rows_kernel<__half, false><<<grid, 256, 0, E->st>>>(E->obj32_ptr, nullptr, nullptr, n, E->n_obj_pad, d, E->d_pad, norms,
                                                                    E->obj_exp, 1, E->obj16.as<__half>(), nullptr);
    else
        convert_rows_kernel<__nv_bfloat16, false><<<grid, 256, 0, E->st>>>(E->obj32_ptr, nullptr, nullptr, n, E->n_obj_pad, d, E->d_pad,
                                                                           norms, 0, 0, E->obj16.as<__nv_bfloat16>(), nullptr);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(E->st));
}

// Shared-memory plan of the fused kernel: subject blocks + object ring + the fixed part of FusedCfg<NW>.
struct TcPlan {
    int kblocks, n_stages, smem_bytes;
    bool ok;
};

template <int NW>
TcPlan plan_fused(int d_pad) {
    TcPlan pl{};
    pl.kblocks = d_pad / tc::KBLK;
    const int fixed = pl.kblocks * tc::BLK_BYTES + tc::FusedCfg<NW>::FIXED_BYTES;
    int stages = (tc::SMEM_LIMIT - fixed) / tc::BLK_BYTES;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    pl.ok = stages >= 2;
    pl.n_stages = stages;
    pl.smem_bytes = fixed + stages * tc::BLK_BYTES;
    return pl;
}

uint32_t make_idesc(bool bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                       // accumulator format: F32
    d |= (bf16 ? 1u : 0u) << 7;         // A format
    d |= (bf16 ? 1u : 0u) << 10;        // B format
    // bits 13/14: no negate; bits 15/16: both operands K-major
    d |= (uint32_t)(tc::TILE_N >> 3) << 17;
    d |= (uint32_t)(256 >> 4) << 24;    // M = 256 across the CTA pair
    return d;
}

int create_impl(b200_rank_engine** out, const void* objects, int32_t dtype, int64_t n_objects, int32_t d, int32_t distance,
                int32_t device, int32_t tc_mode, int32_t flags) {
    if (!out) return fail(B200_E_INVALID, "b200_rank_create: out is NULL");
    *out = nullptr;
    if (n_objects < 0 || d <= 0 || (!objects && n_objects > 0))
        return fail(B200_E_INVALID, "b200_rank_create: bad object matrix (n=%lld, d=%d)", (long long)n_objects, d);
    if (n_objects >= (1ll << 31) - 1) return fail(B200_E_UNSUPPORTED, "b200_rank_create: more than 2^31-2 objects");
    if (distance != B200_DIST_DOT && distance != B200_DIST_COSINE)
        return fail(B200_E_INVALID, "b200_rank_create: distance must be B200_DIST_DOT or B200_DIST_COSINE");
    if (tc_mode < B200_TC_AUTO || tc_mode > B200_TC_OFF) return fail(B200_E_INVALID, "b200_rank_create: bad tc_mode");
    if (dtype < B200_DT_F32 || dtype > B200_DT_BF16) return fail(B200_E_INVALID, "b200_rank_create: bad dtype");
    if (dtype != B200_DT_F32 && !(flags & B200_F_OBJECTS_ON_DEVICE))
        return fail(B200_E_INVALID, "b200_rank_create: 16-bit object factors must be device pointers");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        return fail(B200_E_CUDA, "b200_rank_create: no CUDA device available (the engine has no CPU fallback)");
    if (device < 0 || device >= n_dev) return fail(B200_E_INVALID, "b200_rank_create: device %d out of range", device);
    b200_rank_engine* E = new (std::nothrow) b200_rank_engine();
    if (!E) return fail(B200_E_NOMEM, "b200_rank_create: out of host memory");
    try {
        CK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, device));
        E->device = device;
        E->sm_count = prop.multiProcessorCount;
        E->cc_major = prop.major;
        E->cc_minor = prop.minor;
        snprintf(E->dev_name, sizeof(E->dev_name), "%.127s", prop.name);
        if (prop.major != 10) {
            delete E;
            return fail(B200_E_CUDA, "b200_rank_create: device %d is sm_%d%d; this library contains sm_100a code only",
                        device, prop.major, prop.minor);
        }
        E->distance = distance;
        E->n_obj = n_objects;
        E->d = d;
        E->d_pad = (int)round_up(d, tc::KBLK);
        CK(cudaStreamCreateWithFlags(&E->st, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&E->cs, cudaStreamNonBlocking));
        for (auto& e : E->ev) CK(cudaEventCreate(&e));
        for (auto& e : E->evp) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaMallocHost(&E->h_pinned, 256));
        if ((flags & B200_F_OBJECTS_ON_DEVICE) && dtype == B200_DT_F32) {
            E->obj32_ptr = reinterpret_cast<const float*>(objects);
        } else {
            E->obj32.ensure(sizeof(float) * std::max<int64_t>(n_objects * d, 1));
            if (n_objects > 0) {
                if (dtype == B200_DT_F32) {
                    CK(cudaMemcpyAsync(E->obj32.p, objects, sizeof(float) * n_objects * d, cudaMemcpyHostToDevice, E->st));
                } else {
                    // the caller's stream produced the matrix: order the widening after everything queued on the device
                    CK(cudaDeviceSynchronize());
                    widen16_kernel<<<grid_for(n_objects * d, 256), 256, 0, E->st>>>(objects, dtype == B200_DT_BF16 ? 1 : 0, n_objects * d,
                                                                                   E->obj32.as<float>());
                    CK(cudaGetLastError());
                }
            }
            E->obj32_ptr = E->obj32.as<float>();
        }
        if (tc_mode != B200_TC_OFF && E->d_pad > 1024) tc_mode = B200_TC_OFF;
        if (tc_mode == B200_TC_AUTO && dtype == B200_DT_BF16) tc_mode = B200_TC_BF16;  // bf16 factors: the tensor-core copy is exact
        prepare_objects(E, tc_mode);
        if (E->tc_dtype != B200_TC_OFF) {
            const TcPlan p8 = plan_fused<8>(E->d_pad), p16 = plan_fused<16>(E->d_pad);
            if (!p8.ok || !p16.ok) {
                E->tc_dtype = B200_TC_OFF;
            } else {
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<8, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p8.smem_bytes));
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<8, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p8.smem_bytes));
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<8, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p8.smem_bytes));
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p16.smem_bytes));
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, p16.smem_bytes));
                CK(cudaFuncSetAttribute(tc::fused_topk_kernel<16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, p16.smem_bytes));
            }
        }
        CK(cudaFuncSetAttribute(rescore_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        CK(cudaFuncSetAttribute(rescore_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024));
    } catch (const CudaError& ce) {
        int rc = fail(ce.e == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA, "b200_rank_create: %s failed at line %d: %s",
                      ce.what, ce.line, cudaGetErrorString(ce.e));
        E->free_all();
        delete E;
        return rc;
    }
    *out = E;
    return B200_OK;
}

// Everything one b200_rank_topk call needs, so that the passes below can be plain functions.
struct Call {
    b200_rank_engine* E;
    const b200_rank_query* q;
    b200_rank_stats S;
    cudaStream_t st;
    int64_t n_rows, n_pos;
    int k_out, d;
    // whole-call device views (absolute rows)
    const float* sub32 = nullptr;
    const int64_t* rowmap = nullptr;
    const int64_t* indptr = nullptr;
    const int32_t* indices = nullptr;
    const int32_t* wl = nullptr;
    int32_t *o_ids = nullptr, *o_counts = nullptr;
    float *o_scores = nullptr, *o_bounds = nullptr;
    bool wl_gathered = false;
    bool bf16 = false;
    int nw = 8;  // epilogue warps of the main pass
    size_t n_evt = 0;
    std::vector<int> evt_kind;  // 0 = fused kernel, 1 = selection

    const float* norms() const { return E->distance == B200_DIST_COSINE ? E->obj_norms.as<float>() : nullptr; }

    void time_begin(int kind) {
        if (E->evt.size() < 2 * (n_evt + 1)) {
            cudaEvent_t a, b;
            CK(cudaEventCreate(&a));
            CK(cudaEventCreate(&b));
            E->evt.push_back(a);
            E->evt.push_back(b);
        }
        evt_kind.push_back(kind);
        CK(cudaEventRecord(E->evt[2 * n_evt], st));
    }
    void time_end() {
        CK(cudaEventRecord(E->evt[2 * n_evt + 1], st));
        ++n_evt;
    }
    void collect_times() {  // after the final synchronisation
        for (size_t i = 0; i < n_evt; ++i) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, E->evt[2 * i], E->evt[2 * i + 1]));
            if (evt_kind[i] == 0) {
                S.ms_main += ms;
                S.n_tc_launches++;
            } else {
                S.ms_select += ms;
            }
        }
    }
};

// Exhaustive fp64 passes over `n_sel` rows (rows_dev == nullptr: rows [0, n_sel) relative to the given base pointers).
void run_exact(Call& c, const int32_t* rows_dev, int64_t n_sel, const float* sub32, const int64_t* rowmap, const int64_t* indptr,
               int32_t* o_ids, float* o_scores, int32_t* o_counts, int k_begin, int k_end, bool timed) {
    b200_rank_engine* E = c.E;
    const int64_t tiles_total = (c.n_pos + 31) / 32;
    const int blocks_x = grid_for(n_sel, EX_ROWS);
    int n_splits = (2 * E->sm_count + blocks_x - 1) / blocks_x;
    n_splits = (int)std::max<int64_t>(1, std::min<int64_t>(n_splits, tiles_total / 64));
    n_splits = std::min(n_splits, 1024);
    E->part_scores.ensure(sizeof(float) * (size_t)n_splits * n_sel * LIST_LEN);
    E->part_ids.ensure(sizeof(int32_t) * (size_t)n_splits * n_sel * LIST_LEN);
    for (int k0 = k_begin; k0 < k_end; k0 += 32) {
        const int kp = std::min(32, k_end - k0);
        ExactParams p{};
        p.subjects = sub32;
        p.row_map = rowmap;
        p.rows = rows_dev;
        p.n_sel_dev = nullptr;
        p.n_sel = n_sel;
        p.objects = E->obj32_ptr;
        p.pos2obj = c.wl;
        p.n_pos = c.n_pos;
        p.d = c.d;
        p.obj_norms = c.norms();
        p.indptr = indptr;
        p.indices = c.indices;
        p.id_off = (int32_t)E->id_offset;
        p.k_out = c.k_out;
        p.k0 = k0;
        p.kp = kp;
        p.out_ids = o_ids;
        p.out_scores = o_scores;
        p.out_counts = o_counts;
        p.part_scores = E->part_scores.as<float>();
        p.part_ids = E->part_ids.as<int32_t>();
        p.part_stride_rows = n_sel;
        if (timed) c.time_begin(0);
        exact_topk_kernel<<<dim3(blocks_x, n_splits), EX_THREADS, 0, c.st>>>(p);
        CK(cudaGetLastError());
        if (timed) c.time_end();
        SelectParams sp{};
        sp.in_scores = E->part_scores.as<float>();
        sp.in_ids = E->part_ids.as<int32_t>();
        sp.in_counts = nullptr;
        sp.n_lists = n_splits;
        sp.L = LIST_LEN;
        sp.n_sel = n_sel;
        sp.list_stride_rows = n_sel;
        sp.rows = rows_dev;
        sp.k_out = c.k_out;
        sp.k0 = k0;
        sp.kp = kp;
        sp.out_ids = o_ids;
        sp.out_scores = o_scores;
        sp.out_counts = o_counts;
        merge_select_kernel<<<grid_for(n_sel, SEL_WARPS), SEL_WARPS * 32, 0, c.st>>>(sp);
        CK(cudaGetLastError());
        c.S.n_launches += 2;
    }
    if (timed) c.S.n_splits = n_splits;
}

struct TcPass {
    const int32_t* rows_dev = nullptr;  // nullptr: rows [0, n_sel) of the base pointers below
    int64_t n_sel = 0;
    // base pointers: the chunk's slice (rows_dev == nullptr) or the whole call (rows_dev = absolute rows)
    const float* sub32 = nullptr;
    const int64_t* rowmap = nullptr;
    const int64_t* indptr = nullptr;
    int32_t* o_ids = nullptr;
    float* o_scores = nullptr;
    int32_t* o_counts = nullptr;
    float* o_bounds = nullptr;  // shared-threshold mode: bounds out, no verdict
    int nw = 8;                 // epilogue warps
    int kc = 12;                // K' per list
    int k0 = 0, kp = 0;         // this pass produces entries [k0, k0 + kp)
    bool wide = false;          // single-pass wide mode (frozen threshold + global append)
    bool peers = false;         // share thresholds with the other ranks
    int64_t row0 = 0;           // absolute row of batch row 0 (failure list entries, peer arrays)
    int32_t* fb_list = nullptr;
    int32_t* fb_count = nullptr;
    bool main = false;          // reported in the statistics as the main pass
};

// One tensor-core candidate pass + fp64 re-score + certificate.  Rows whose certificate fails are appended to `fb_list`.
void run_tc(Call& c, const TcPass& t) {
    b200_rank_engine* E = c.E;
    cudaStream_t st = c.st;
    const bool bf16 = c.bf16;
    const int d = c.d;
    const int nlist = t.nw / 4;
    const int slots = 64 / nlist;
    const TcPlan pl = t.nw == 8 ? plan_fused<8>(E->d_pad) : plan_fused<16>(E->d_pad);
    const int64_t rows_pad = round_up(t.n_sel, 2 * tc::TILE_M);
    // subjects -> 16-bit, per-row power-of-two scale
    E->sub16.ensure((size_t)rows_pad * E->d_pad * 2);
    E->row_exp.ensure(sizeof(int32_t) * rows_pad);
    {
        const int grid = grid_for(rows_pad * 32, 256);
        if (!bf16)
            convert_rows_kernel<__half, true><<<grid, 256, 0, st>>>(t.sub32, t.rowmap, t.rows_dev, t.n_sel, rows_pad, d, E->d_pad, nullptr, 0, 1,
                                                                    E->sub16.as<__half>(), E->row_exp.as<int32_t>());
        else
            convert_rows_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(t.sub32, t.rowmap, t.rows_dev, t.n_sel, rows_pad, d, E->d_pad, nullptr,
                                                                           0, 0, E->sub16.as<__nv_bfloat16>(), E->row_exp.as<int32_t>());
        CK(cudaGetLastError());
        c.S.n_launches++;
    }
    // objects: resident 16-bit copy, or a whitelist gather of it
    const void* obj_base = E->obj16.p;
    int64_t obj_rows = E->n_obj_pad;
    if (c.wl) {
        const int64_t npad = round_up(c.n_pos, tc::HALF_N);
        if (!c.wl_gathered) {  // shared by every chunk / pass / re-rank of the call
            c.wl_gathered = true;
            E->obj16_wl.ensure((size_t)npad * E->d_pad * 2);
            const int chunks = E->d_pad * 2 / 16;
            gather_rows16_kernel<<<grid_for(npad * chunks, 256), 256, 0, st>>>(E->obj16.as<uint4>(), c.wl, c.n_pos, npad, chunks,
                                                                             E->obj16_wl.as<uint4>());
            CK(cudaGetLastError());
            c.S.n_launches++;
        }
        obj_base = E->obj16_wl.p;
        obj_rows = npad;
    }
    CUtensorMap tm_obj, tm_sub;
    if (!make_tensor_map(&tm_obj, obj_base, obj_rows, E->d_pad, bf16) || !make_tensor_map(&tm_sub, E->sub16.p, rows_pad, E->d_pad, bf16))
        throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled", __LINE__};

    tc::TcParams tp{};
    tp.kblocks = pl.kblocks;
    tp.n_stages = pl.n_stages;
    tp.k_cand = std::min(t.kc, slots);
    tp.n_rows = t.n_sel;
    tp.n_pos = c.n_pos;
    tp.n_row_tiles = (int)(rows_pad / (2 * tc::TILE_M));
    tp.n_obj_tiles = (int)((c.n_pos + tc::TILE_N - 1) / tc::TILE_N);
    // object splits: fill the machine when there are few row tiles, even out the last wave otherwise
    int best_splits = 1;
    double best_eff = -1.0;
    const int max_splits = t.wide ? 1 : std::max(1, std::min(16, tp.n_obj_tiles * 2 / 32));
    const int n_units = E->sm_count / 2;  // CTA pairs working concurrently
    for (int s = 1; s <= max_splits; ++s) {
        const double work = (double)tp.n_row_tiles * s;
        const double waves = std::ceil(work / n_units);
        const double eff = work / (waves * n_units) - 0.01 * (s - 1);
        if (eff > best_eff + 1e-9) {
            best_eff = eff;
            best_splits = s;
        }
    }
    {
        const int forced = env_int("B200_TC_SPLITS", 0);  // tuning / test hook
        if (forced >= 1 && forced <= max_splits) best_splits = forced;
    }
    tp.n_splits = best_splits;
    tp.tiles_per_split = (tp.n_obj_tiles + best_splits - 1) / best_splits;
    tp.idesc = make_idesc(bf16);
    tp.pos2obj = c.wl;
    tp.indptr = t.indptr;
    tp.indices = c.indices;
    tp.row_ids = t.rows_dev;
    if (t.k0 > 0) {  // objects returned by earlier passes are excluded like viewed ones
        tp.excl = E->excl.as<int32_t>();
        tp.excl_stride = c.k_out;
        tp.excl_n = t.k0;
    }
    tp.id_off = (int32_t)E->id_offset;
    const int n_lists = best_splits * nlist;
    // wide mode: the lists hold ~T = 1.35 k + 40 candidates per row -- the threshold frozen after a fraction q of the stream
    // is about the (lists x K' - 6)-th best of that fraction, i.e. rank ~ (lists x K' - 6) / q overall
    int cand_stride = 32;
    tp.phase1_tiles = 0x7fffffff;
    if (t.wide) {
        const double T = env_int("B200_WIDE_T", (int)(1.35 * t.kp + 40));
        const double rank_frozen = nlist * tp.k_cand - 6;
        const double qf = std::min(1.0, rank_frozen / T);
        tp.phase1_tiles = std::max(1, (int)std::ceil(qf * tp.tiles_per_split));
        cand_stride = (int)round_up((int64_t)(T / nlist * 1.5 + 32), 8);
        cand_stride = std::min(cand_stride, WIDE_MAX / nlist);
    }
    tp.cand_stride = cand_stride;
    E->cand_scores.ensure(sizeof(float) * (size_t)n_lists * rows_pad * cand_stride);
    E->cand_ids.ensure(sizeof(int32_t) * (size_t)n_lists * rows_pad * cand_stride);
    E->cand_counts.ensure(sizeof(int32_t) * (size_t)n_lists * rows_pad);
    E->cand_thr.ensure(sizeof(float) * (size_t)n_lists * rows_pad);
    tp.cand_scores = E->cand_scores.as<float>();
    tp.cand_ids = E->cand_ids.as<int32_t>();
    tp.cand_counts = E->cand_counts.as<int32_t>();
    tp.cand_thr = E->cand_thr.as<float>();
    tp.rows_pad = rows_pad;
    tp.debug_mode = env_int("B200_TC_DEBUG", 0);  // measurement hook, results are invalid
    if (t.peers) {
        tp.n_peers = E->n_peers;
        tp.peer_epoch = c.q->peer_epoch;
        tp.peer_exp = E->obj_exp;
        tp.peer_row0 = t.row0;
        tp.peer_pub = E->peer_pub.as<unsigned long long>();
        for (int i = 0; i < E->n_peers; ++i) tp.peer_in[i] = reinterpret_cast<const unsigned long long*>(E->peer_in[i]);
    }
    if (t.main) {
        c.S.n_splits = best_splits;
        c.S.k_cand = tp.k_cand;
        c.S.epi_warps = t.nw;
        c.S.wide = t.wide ? 1 : 0;
    }
    const int n_work = tp.n_row_tiles * tp.n_splits;
    if (env_int("B200_TC_CAROUSEL", 1) != 0) {  // 0: every work item starts at its first object tile
        const int n_pairs_run = std::min(n_work, n_units);
        const int per_pair = (n_work + n_pairs_run - 1) / n_pairs_run;
        const int64_t n_ints = (int64_t)best_splits + (int64_t)n_pairs_run * per_pair;
        E->carousel.ensure(sizeof(int32_t) * n_ints);
        carousel_init_kernel<<<grid_for(n_ints, 256), 256, 0, st>>>(E->carousel.as<int32_t>(), best_splits, tp.tiles_per_split, n_ints);
        CK(cudaGetLastError());
        c.S.n_launches++;
        tp.front = E->carousel.as<int32_t>();
        tp.starts = tp.front + best_splits;
        tp.starts_stride = per_pair;
    }
    const int grid = 2 * std::min(n_work, n_units);
    c.time_begin(0);
    const bool use_peers = t.peers && tp.n_peers > 0;  // (wide mode and threshold sharing never meet: sharing needs k <= 24)
#define B200_LAUNCH(NW_, WIDE_, PEERS_) \
    tc::fused_topk_kernel<NW_, WIDE_, PEERS_><<<grid, tc::FusedCfg<NW_>::THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp)
    if (t.nw == 8) {
        if (t.wide) B200_LAUNCH(8, true, false);
        else if (use_peers) B200_LAUNCH(8, false, true);
        else B200_LAUNCH(8, false, false);
    } else {
        if (t.wide) B200_LAUNCH(16, true, false);
        else if (use_peers) B200_LAUNCH(16, false, true);
        else B200_LAUNCH(16, false, false);
    }
#undef B200_LAUNCH
    CK(cudaGetLastError());
    c.time_end();
    c.S.n_launches++;

    // fp64 re-score of the candidates + certificate
    SelectParams sp{};
    sp.in_scores = tp.cand_scores;
    sp.in_ids = tp.cand_ids;
    sp.in_counts = tp.cand_counts;
    sp.in_thr = tp.cand_thr;
    sp.n_lists = n_lists;
    sp.L = cand_stride;
    sp.n_sel = t.n_sel;
    sp.list_stride_rows = rows_pad;
    sp.rows = t.rows_dev;
    sp.k_out = c.k_out;
    sp.k0 = t.k0;
    sp.kp = t.kp;
    sp.out_ids = t.o_ids;
    sp.out_scores = t.o_scores;
    sp.out_counts = t.o_counts;
    sp.subjects = t.sub32;
    sp.row_map = t.rowmap;
    sp.objects = E->obj32_ptr;
    sp.obj_norms = c.norms();
    sp.d = d;
    sp.row_exp = E->row_exp.as<int32_t>();
    sp.obj_exp = E->obj_exp;
    // |approx - exact| <= eps_rel * |u|_2 * max |i|_2: rounding of both operands (rho each; none for factors that are exact
    // in the tensor-core type) plus a generous bound on the tensor-core accumulation
    const double rho = bf16 ? 0.001953125 /*2^-9*/ : 0.00048828125 /*2^-11*/;
    sp.eps_rel = (float)(2.0 * rho + rho * rho + (double)E->d_pad * 4.76837158e-7 /*2^-21*/ + std::sqrt((double)d) * 1.4551915e-11 /*2^-36*/);
    sp.max_obj_norm = E->max_obj_norm;
    sp.fb_count = t.fb_count;
    sp.fb_rows = t.fb_list;
    sp.fb_row0 = t.rows_dev ? 0 : t.row0;
    sp.out_bounds = t.o_bounds;
    c.time_begin(1);
    if (t.wide) {
        rescore_wide_kernel<<<(unsigned)t.n_sel, WIDE_THREADS, (size_t)d * sizeof(float), st>>>(sp);
    } else {
        const size_t sel_smem = (size_t)SEL_WARPS * d * sizeof(float);
        rescore_select_kernel<<<grid_for(t.n_sel, SEL_WARPS), SEL_WARPS * 32, sel_smem, st>>>(sp);
    }
    CK(cudaGetLastError());
    c.time_end();
    c.S.n_launches++;
}

// Sparse subjects (EASE): SpMM score rows for bounded row chunks + streaming top-k (sparse.cuh).
void run_sparse(Call& c, const int64_t* sp_indptr, const int32_t* sp_indices, const float* sp_data, int64_t nr, const int64_t* f_indptr,
                int32_t* o_ids, float* o_scores, int32_t* o_counts) {
    b200_rank_engine* E = c.E;
    cudaStream_t st = c.st;
    if (!E->objT.p && E->n_obj > 0) {  // transposed master copy, built once
        E->objT.ensure(sizeof(float) * (size_t)E->n_obj * E->d);
        transpose_kernel<<<dim3((unsigned)grid_for(E->n_obj, 32), (unsigned)grid_for(E->d, 32)), dim3(32, 8), 0, st>>>(E->obj32_ptr, E->n_obj, E->d,
                                                                                                                  E->objT.as<float>());
        CK(cudaGetLastError());
        c.S.n_launches++;
    }
    const int64_t rows_max = std::max<int64_t>(1, std::min<int64_t>(nr, ((int64_t)1 << 30) / std::max<int64_t>(4 * c.n_pos, 1)));
    E->sp_scores.ensure(sizeof(float) * (size_t)rows_max * c.n_pos);
    for (int64_t b0 = 0; b0 < nr; b0 += rows_max) {
        const int64_t nb = std::min(rows_max, nr - b0);
        c.time_begin(0);
        sparse_scores_kernel<<<dim3((unsigned)nb, (unsigned)grid_for(c.n_pos, SP_BLOCK_COLS)), SP_THREADS, 0, st>>>(
            sp_indptr + b0, sp_indices, sp_data, E->objT.as<float>(), E->n_obj, E->d, c.wl, c.n_pos, E->sp_scores.as<float>());
        CK(cudaGetLastError());
        c.time_end();
        c.S.n_launches++;
        for (int k0 = 0; k0 < c.k_out; k0 += 32) {
            c.time_begin(1);
            scores_topk_kernel<<<grid_for(nb * 32, 256), 256, 0, st>>>(E->sp_scores.as<float>(), nb, c.n_pos, c.wl, f_indptr ? f_indptr + b0 : nullptr,
                                                                      c.indices, (int32_t)E->id_offset, c.k_out, k0, std::min(32, c.k_out - k0),
                                                                      o_ids + b0 * c.k_out, o_scores + b0 * c.k_out, o_counts + b0);
            CK(cudaGetLastError());
            c.time_end();
            c.S.n_launches++;
        }
    }
}

// Dense subjects with k > 128: one exhaustive scoring of bounded row chunks into HBM + k / 32 streaming selection passes.
void run_dense_large_k(Call& c, const float* sub32, const int64_t* rowmap, const int64_t* f_indptr, int64_t nr, int32_t* o_ids,
                       float* o_scores, int32_t* o_counts) {
    b200_rank_engine* E = c.E;
    cudaStream_t st = c.st;
    const int64_t rows_max = std::max<int64_t>(32, std::min<int64_t>(nr, ((int64_t)1 << 30) / std::max<int64_t>(4 * c.n_pos, 1)) / 32 * 32);
    E->sp_scores.ensure(sizeof(float) * (size_t)rows_max * c.n_pos);
    for (int64_t b0 = 0; b0 < nr; b0 += rows_max) {
        const int64_t nb = std::min(rows_max, nr - b0);
        const int blocks_x = grid_for(nb, 32);
        const int64_t tiles_total = (c.n_pos + 31) / 32;
        const int splits = (int)std::max<int64_t>(1, std::min<int64_t>((4 * E->sm_count + blocks_x - 1) / blocks_x, tiles_total));
        c.time_begin(0);
        dense_scores_kernel<<<dim3((unsigned)blocks_x, (unsigned)splits), 256, 0, st>>>(
            rowmap ? sub32 : sub32 + b0 * c.d, rowmap ? rowmap + b0 : nullptr, nb, E->obj32_ptr, c.wl, c.n_pos, c.d, c.norms(),
            E->sp_scores.as<float>());
        CK(cudaGetLastError());
        c.time_end();
        c.S.n_launches++;
        for (int k0 = 0; k0 < c.k_out; k0 += 32) {
            c.time_begin(1);
            scores_topk_kernel<<<grid_for(nb * 32, 256), 256, 0, st>>>(E->sp_scores.as<float>(), nb, c.n_pos, c.wl, f_indptr ? f_indptr + b0 : nullptr,
                                                                      c.indices, (int32_t)E->id_offset, c.k_out, k0, std::min(32, c.k_out - k0),
                                                                      o_ids + b0 * c.k_out, o_scores + b0 * c.k_out, o_counts + b0);
            CK(cudaGetLastError());
            c.time_end();
            c.S.n_launches++;
        }
    }
}

int32_t read_counter(Call& c, const int32_t* dev) {
    CK(cudaMemcpyAsync(c.E->h_pinned, dev, sizeof(int32_t), cudaMemcpyDeviceToHost, c.st));
    CK(cudaStreamSynchronize(c.st));
    return c.E->h_pinned[0];
}

}  // namespace

extern "C" {

const char* b200_rank_last_error(void) { return g_last_error.c_str(); }
int b200_rank_abi_version(void) { return B200_RANK_ABI_VERSION; }

int b200_rank_create(b200_rank_engine** out, const float* objects, int64_t n_objects, int32_t d, int32_t distance,
                     int32_t device, int32_t tc_mode, int32_t flags) {
    return create_impl(out, objects, B200_DT_F32, n_objects, d, distance, device, tc_mode, flags);
}

int b200_rank_create_ex(b200_rank_engine** out, const void* objects, int32_t dtype, int64_t n_objects, int32_t d, int32_t distance,
                        int32_t device, int32_t tc_mode, int32_t flags) {
    return create_impl(out, objects, dtype, n_objects, d, distance, device, tc_mode, flags);
}

int b200_rank_destroy(b200_rank_engine* E) {
    if (!E) return B200_OK;
    cudaSetDevice(E->device);
    if (E->st) cudaStreamSynchronize(E->st);
    E->free_all();
    delete E;
    return B200_OK;
}

int b200_rank_get_info(b200_rank_engine* E, b200_rank_info* info) {
    if (!E || !info) return fail(B200_E_INVALID, "b200_rank_get_info: NULL argument");
    memset(info, 0, sizeof(*info));
    info->abi_version = B200_RANK_ABI_VERSION;
    info->device = E->device;
    info->sm_count = E->sm_count;
    info->cc_major = E->cc_major;
    info->cc_minor = E->cc_minor;
    info->tc_dtype = E->tc_dtype;
    info->n_objects = E->n_obj;
    info->d = E->d;
    info->d_pad = E->d_pad;
    info->hbm_bytes = (int64_t)E->hbm_bytes();
    snprintf(info->device_name, sizeof(info->device_name), "%s", E->dev_name);
    return B200_OK;
}

int b200_rank_set_subjects(b200_rank_engine* E, const float* subjects, int64_t n_subjects, int32_t on_device) {
    if (!E) return fail(B200_E_INVALID, "b200_rank_set_subjects: engine is NULL");
    if (n_subjects < 0 || (!subjects && n_subjects > 0)) return fail(B200_E_INVALID, "b200_rank_set_subjects: bad matrix");
    std::lock_guard<std::mutex> lock(E->mu);
    try {
        CK(cudaSetDevice(E->device));
        if (on_device) {
            E->sub32_res_ptr = subjects;
        } else {
            E->sub32_res.ensure(sizeof(float) * std::max<int64_t>(n_subjects * E->d, 1));
            if (n_subjects > 0)
                CK(cudaMemcpyAsync(E->sub32_res.p, subjects, sizeof(float) * n_subjects * E->d, cudaMemcpyHostToDevice, E->st));
            CK(cudaStreamSynchronize(E->st));
            E->sub32_res_ptr = E->sub32_res.as<float>();
        }
        E->n_sub_res = n_subjects;
    } catch (const CudaError& ce) {
        return fail(ce.e == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA, "b200_rank_set_subjects: %s failed: %s", ce.what,
                    cudaGetErrorString(ce.e));
    }
    return B200_OK;
}

int b200_rank_set_id_offset(b200_rank_engine* E, int64_t offset) {
    if (!E) return fail(B200_E_INVALID, "b200_rank_set_id_offset: engine is NULL");
    if (offset < 0 || offset + E->n_obj >= (1ll << 31) - 1)
        return fail(B200_E_INVALID, "b200_rank_set_id_offset: offset + n_objects must stay below 2^31-1");
    std::lock_guard<std::mutex> lock(E->mu);
    E->id_offset = offset;
    return B200_OK;
}

int b200_rank_peer_export(b200_rank_engine* E, int64_t max_rows, void* handle_out) {
    if (!E || !handle_out || max_rows <= 0) return fail(B200_E_INVALID, "b200_rank_peer_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI hands 64-byte handles around");
    std::lock_guard<std::mutex> lock(E->mu);
    try {
        CK(cudaSetDevice(E->device));
        if (E->peer_pub.p) return fail(B200_E_INVALID, "b200_rank_peer_export: already exported (the peers hold the old handle)");
        E->peer_pub.ensure(sizeof(unsigned long long) * max_rows);
        CK(cudaMemset(E->peer_pub.p, 0, E->peer_pub.cap));
        E->peer_rows = max_rows;
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, E->peer_pub.p));
        memcpy(handle_out, &h, sizeof(h));
    } catch (const CudaError& ce) {
        return fail(B200_E_CUDA, "b200_rank_peer_export: %s failed: %s", ce.what, cudaGetErrorString(ce.e));
    }
    return B200_OK;
}

int b200_rank_peer_import(b200_rank_engine* E, int32_t n_ranks, int32_t self, const void* handles) {
    if (!E || !handles || n_ranks < 1 || self < 0 || self >= n_ranks) return fail(B200_E_INVALID, "b200_rank_peer_import: bad arguments");
    if (n_ranks - 1 > tc::MAX_PEERS) return fail(B200_E_UNSUPPORTED, "b200_rank_peer_import: at most %d ranks", tc::MAX_PEERS + 1);
    std::lock_guard<std::mutex> lock(E->mu);
    if (!E->peer_pub.p) return fail(B200_E_INVALID, "b200_rank_peer_import: call b200_rank_peer_export first");
    try {
        CK(cudaSetDevice(E->device));
        int n = 0;
        for (int r = 0; r < n_ranks; ++r) {
            if (r == self) continue;
            cudaIpcMemHandle_t h;
            memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)r * 64, 64);
            void* ptr = nullptr;
            CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
            E->peer_in[n++] = ptr;
        }
        E->n_peers = n;
    } catch (const CudaError& ce) {
        return fail(B200_E_CUDA, "b200_rank_peer_import: %s failed: %s", ce.what, cudaGetErrorString(ce.e));
    }
    return B200_OK;
}

int b200_rank_topk(b200_rank_engine* E, const b200_rank_query* q, b200_rank_stats* stats) {
    if (!E || !q) return fail(B200_E_INVALID, "b200_rank_topk: NULL argument");
    if (q->n_rows < 0) return fail(B200_E_INVALID, "b200_rank_topk: n_rows < 0");
    if (q->k <= 0) return fail(B200_E_INVALID, "b200_rank_topk: k must be positive");
    const bool sparse_sub = q->sub_indptr != nullptr;
    if (!sparse_sub && !q->subjects && !q->subject_ids) return fail(B200_E_INVALID, "b200_rank_topk: neither subjects nor subject_ids given");
    if (sparse_sub && (q->subjects || q->subject_ids)) return fail(B200_E_INVALID, "b200_rank_topk: sparse subjects exclude subjects / subject_ids");
    if (sparse_sub && E->distance != B200_DIST_DOT)
        return fail(B200_E_INVALID, "b200_rank_topk: sparse subjects need B200_DIST_DOT (rank_implicit.py:66-67)");
    if (!sparse_sub && !q->subjects && !E->sub32_res_ptr)
        return fail(B200_E_INVALID, "b200_rank_topk: subject_ids given but b200_rank_set_subjects was never called");
    if (q->subjects && q->subject_ids && q->n_subjects_total <= 0)
        return fail(B200_E_INVALID, "b200_rank_topk: subjects + subject_ids need n_subjects_total");
    if (q->whitelist && q->n_whitelist < 0) return fail(B200_E_INVALID, "b200_rank_topk: n_whitelist < 0");
    if (q->n_rows > 0 && (!q->out_ids || !q->out_scores || !q->out_counts))
        return fail(B200_E_INVALID, "b200_rank_topk: output pointers are NULL");
    if (q->n_rows >= (1ll << 31) - 64) return fail(B200_E_UNSUPPORTED, "b200_rank_topk: more than 2^31 rows per call");
    if ((q->flags & B200_Q_FORCE_EXACT) && (q->flags & B200_Q_FORCE_TC))
        return fail(B200_E_INVALID, "b200_rank_topk: FORCE_EXACT and FORCE_TC are exclusive");
    const bool in_dev = q->flags & B200_Q_INPUTS_ON_DEVICE;
    const bool out_dev = q->flags & B200_Q_OUTPUTS_ON_DEVICE;
    const bool shared = q->flags & B200_Q_SHARED_THRESHOLDS;
    if (q->subject_dtype != B200_DT_F32 && !(in_dev && q->subjects && !q->subject_ids))
        return fail(B200_E_INVALID, "b200_rank_topk: 16-bit subjects must be a device matrix in batch order");
    if (q->subject_dtype < B200_DT_F32 || q->subject_dtype > B200_DT_BF16) return fail(B200_E_INVALID, "b200_rank_topk: bad subject_dtype");
    if (shared && (!q->out_bounds || q->peer_epoch == 0))
        return fail(B200_E_INVALID, "b200_rank_topk: B200_Q_SHARED_THRESHOLDS needs out_bounds and peer_epoch >= 1");
    if (shared && sparse_sub) return fail(B200_E_UNSUPPORTED, "b200_rank_topk: sparse subjects cannot share thresholds");

    std::lock_guard<std::mutex> lock(E->mu);
    Call c{};
    c.E = E;
    c.q = q;
    memset(&c.S, 0, sizeof(c.S));
    b200_rank_stats& S = c.S;
    const int64_t n_rows = q->n_rows;
    const int64_t n_pos = q->whitelist ? q->n_whitelist : E->n_obj;
    const int k_out = (int)std::min<int64_t>(q->k, n_pos);
    S.k_out = k_out;
    const int d = E->d;
    c.n_rows = n_rows;
    c.n_pos = n_pos;
    c.k_out = k_out;
    c.d = d;
    if (n_rows == 0 || k_out <= 0) {
        if (stats) *stats = S;
        return B200_OK;
    }
    if (shared && E->n_peers > 0 && n_rows > E->peer_rows)
        return fail(B200_E_INVALID, "b200_rank_topk: %lld rows exceed the %lld exported for threshold sharing", (long long)n_rows,
                    (long long)E->peer_rows);
    try {
        CK(cudaSetDevice(E->device));
        cudaStream_t st = E->st;
        c.st = st;
        cudaStream_t user = reinterpret_cast<cudaStream_t>(q->stream);
        // device pointers + NULL stream = CUDA's (legacy) default stream, like every CUDA API: producers / consumers of the
        // buffers on that stream are ordered against the engine stream (torch's current stream is the default stream unless
        // the caller switched it: without this a collective reading the outputs could overlap the next call's kernels)
        if (!user && (in_dev || out_dev)) user = cudaStreamLegacy;
        if (user && (in_dev || out_dev)) {
            CK(cudaEventRecord(E->ev[6], user));
            CK(cudaStreamWaitEvent(st, E->ev[6], 0));
        }
        CK(cudaEventRecord(E->ev[0], st));

        // ---------------- stage inputs
        // Host inputs of a large call are staged in row chunks on a second stream: the copy of chunk c+1 (subject rows / ids,
        // its slice of the CSR filter) and the copy-back of chunk c-1 run while chunk c is being ranked.  Buffers are
        // full-size and addressed by absolute row / nnz offsets, so the kernels see the same layout with or without chunking.
        auto stage = [&](DevBuf& buf, const void* src, size_t bytes) -> const void* {  // un-chunked items, main stream
            if (in_dev) return src;
            buf.ensure(std::max<size_t>(bytes, 16));
            if (bytes) CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
            S.h2d_bytes += (int64_t)bytes;
            return buf.p;
        };
        const bool chunk_subjects = q->subjects && !q->subject_ids && !in_dev;  // subject rows arrive in batch order
        const int64_t* sp_indptr = nullptr;
        const int32_t* sp_indices = nullptr;
        const float* sp_data = nullptr;
        if (sparse_sub) {
            int64_t nnz = 0;
            if (in_dev) {
                CK(cudaMemcpyAsync(E->h_pinned, q->sub_indptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                memcpy(&nnz, E->h_pinned, sizeof(int64_t));
            } else {
                nnz = q->sub_indptr[n_rows];
            }
            if (nnz < 0 || (nnz > 0 && (!q->sub_indices || !q->sub_data))) return fail(B200_E_INVALID, "b200_rank_topk: bad sparse subjects");
            sp_indptr = (const int64_t*)stage(E->sp_indptr, q->sub_indptr, sizeof(int64_t) * (n_rows + 1));
            sp_indices = (const int32_t*)stage(E->sp_indices, q->sub_indices, sizeof(int32_t) * nnz);
            sp_data = (const float*)stage(E->sp_data, q->sub_data, sizeof(float) * nnz);
        } else if (q->subjects) {
            const int64_t rows_in = q->subject_ids ? q->n_subjects_total : n_rows;
            if (q->subject_dtype != B200_DT_F32) {
                E->sub32.ensure(sizeof(float) * rows_in * d);
                widen16_kernel<<<grid_for(rows_in * d, 256), 256, 0, st>>>(q->subjects, q->subject_dtype == B200_DT_BF16 ? 1 : 0, rows_in * d,
                                                                           E->sub32.as<float>());
                CK(cudaGetLastError());
                S.n_launches++;
                c.sub32 = E->sub32.as<float>();
            } else if (chunk_subjects) {
                E->sub32.ensure(std::max<size_t>(sizeof(float) * rows_in * d, 16));
                c.sub32 = E->sub32.as<float>();
            } else {
                c.sub32 = (const float*)stage(E->sub32, q->subjects, sizeof(float) * rows_in * d);
            }
        } else {
            c.sub32 = E->sub32_res_ptr;
        }
        if (q->subject_ids) {
            if (in_dev) {
                c.rowmap = q->subject_ids;
            } else {
                E->rowmap.ensure(std::max<size_t>(sizeof(int64_t) * n_rows, 16));
                c.rowmap = E->rowmap.as<int64_t>();
            }
        }
        if (q->csr_indptr) {
            int64_t nnz = 0;
            if (in_dev) {
                CK(cudaMemcpyAsync(E->h_pinned, q->csr_indptr + n_rows, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                memcpy(&nnz, E->h_pinned, sizeof(int64_t));
            } else {
                nnz = q->csr_indptr[n_rows];
            }
            if (nnz < 0) return fail(B200_E_INVALID, "b200_rank_topk: csr_indptr[n_rows] < 0");
            if (nnz > 0 && !q->csr_indices) return fail(B200_E_INVALID, "b200_rank_topk: csr_indices is NULL");
            if (in_dev) {
                c.indptr = q->csr_indptr;
                c.indices = q->csr_indices;
            } else {
                E->indptr.ensure(sizeof(int64_t) * (n_rows + 1));
                E->indices.ensure(std::max<size_t>(sizeof(int32_t) * nnz, 16));
                c.indptr = E->indptr.as<int64_t>();
                c.indices = E->indices.as<int32_t>();
            }
            if (nnz == 0) c.indptr = nullptr;  // an all-empty filter is no filter (cf. rank_implicit.py:169-173)
        }
        if (q->whitelist) c.wl = (const int32_t*)stage(E->wl, q->whitelist, sizeof(int32_t) * n_pos);
        // host -> device copy of the chunked inputs of rows [r0, r1) on stream `s`
        auto stage_rows = [&](int64_t r0, int64_t r1, cudaStream_t s) {
            if (in_dev) return;
            size_t bytes = 0;
            auto h2d = [&](void* dst, const void* src, size_t n) {
                if (n) CK(cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, s));
                bytes += n;
            };
            if (chunk_subjects) h2d(E->sub32.as<float>() + r0 * d, q->subjects + r0 * d, sizeof(float) * (r1 - r0) * d);
            if (q->subject_ids) h2d(E->rowmap.as<int64_t>() + r0, q->subject_ids + r0, sizeof(int64_t) * (r1 - r0));
            if (c.indptr) {
                h2d(E->indptr.as<int64_t>() + r0, q->csr_indptr + r0, sizeof(int64_t) * (r1 - r0 + 1));
                const int64_t z0 = q->csr_indptr[r0], z1 = q->csr_indptr[r1];
                if (z1 < z0) throw CudaError{cudaErrorInvalidValue, "csr_indptr must be non-decreasing", __LINE__};
                h2d(E->indices.as<int32_t>() + z0, q->csr_indices + z0, sizeof(int32_t) * (z1 - z0));
            }
            S.h2d_bytes += (int64_t)bytes;
        };

        // ---------------- outputs
        if (out_dev) {
            c.o_ids = q->out_ids;
            c.o_scores = q->out_scores;
            c.o_counts = q->out_counts;
            c.o_bounds = shared ? q->out_bounds : nullptr;
        } else {
            E->out_ids.ensure(sizeof(int32_t) * n_rows * k_out);
            E->out_scores.ensure(sizeof(float) * n_rows * k_out);
            E->out_counts.ensure(sizeof(int32_t) * n_rows);
            c.o_ids = E->out_ids.as<int32_t>();
            c.o_scores = E->out_scores.as<float>();
            c.o_counts = E->out_counts.as<int32_t>();
            if (shared) {
                E->out_bounds.ensure(sizeof(float) * n_rows);
                c.o_bounds = E->out_bounds.as<float>();
            }
        }
        // ---------------- path choice
        // Candidates kept per list by the tensor-core pass (K' >= k / lists; the surplus is the certificate's safety margin).
        // A row has 2 (8 epilogue warps) or 4 (16) lists, one per column group of the tile stream, so a small surplus per
        // list already gives ~2k candidates; rows where (nearly) all of the top-k fall into one column group fail the
        // certificate and take the second-chance pass.  Inserts, the dominant epilogue cost, scale with K'.
        c.bf16 = E->tc_dtype == B200_TC_BF16;
        const bool wide = k_out > 24 && k_out <= 128 && env_int("B200_WIDE", 1) != 0;
        // 16 epilogue warps (B200_EPI_WARPS=16) measured 5 % slower at N = 1M, 5 % faster on a 125 K-object shard, equal at
        // d = 256 (profiles/r02_ab_fused.txt): opt-in.  The wide mode always runs the 8-warp geometry: four lists per row
        // freeze at a weaker, noisier rank and 10 % of the rows miss their candidate count.
        c.nw = (!wide && env_int("B200_EPI_WARPS", 8) == 16) ? 16 : 8;
        int k_cand = 0;
        if (k_out <= 24) {
            if (c.nw == 16) {
                // four lists per row: a list may be SHORTER than k (the certificate only needs the k-th exact score above every
                // list's threshold); rows whose top-k crowd into one column quarter take the second-chance pass
                k_cand = std::min(16, (k_out <= 10 ? 8 : k_out <= 16 ? 12 : 16) + (c.bf16 ? 2 : 0));
            } else {
                const int surplus = c.bf16 ? std::max(6, k_out / 2) : std::max(2, k_out / 4);
                k_cand = std::min(32, k_out + surplus);
            }
        } else if (k_out <= 128) {
            k_cand = wide ? 24 : (c.bf16 ? 30 : 25);  // wide: adaptive lists of phase 1;  else passes of 20
        }
        if (shared && E->n_peers > 0 && k_out <= 24) {
            // Shared thresholds: the pruning bound of a row is the MAXIMUM over all L = ranks x lists list minima, i.e. the
            // largest K'-th best of L samples of N/L objects -- about global rank L K' - c_L L sqrt(K') (c_L = expected maximum
            // of L standard normals).  The certificate needs that rank to stay above k plus a margin; everything beyond is
            // wasted insertions (K' = 12 on 8 ranks sits near rank 100, K' = 6 near rank 27).
            const int L = (E->n_peers + 1) * (c.nw / 4);
            const double cL = L <= 2 ? 0.56 : L <= 4 ? 1.03 : L <= 8 ? 1.42 : L <= 16 ? 1.77 : L <= 32 ? 2.07 : 2.33;
            const double target = k_out + std::max(12.0, 0.6 * k_out) + (c.bf16 ? 20.0 : 0.0);
            int kc = 4;
            while (kc < 32 && L * kc - cL * L * std::sqrt((double)kc) < target) ++kc;
            k_cand = std::min(kc, c.nw == 16 ? 16 : 32);
        }
        {
            const int forced = env_int("B200_TC_KCAND", 0);  // tuning hook
            if (forced >= 4 && forced <= 32 && (forced >= k_out || c.nw == 16 || wide || shared)) k_cand = forced;
        }
        bool use_tc = !sparse_sub && E->tc_dtype != B200_TC_OFF && k_cand > 0 && !(q->flags & B200_Q_FORCE_EXACT) && n_pos >= (int64_t)k_cand * 4;
        if (use_tc && !(q->flags & B200_Q_FORCE_TC)) {
            // tiny problems are cheaper (and exercised) on the exhaustive kernel
            if ((double)n_rows * (double)n_pos < 4.0e6) use_tc = false;
        }
        if (use_tc && (size_t)SEL_WARPS * d * sizeof(float) > 64 * 1024)
            return fail(B200_E_UNSUPPORTED, "b200_rank_topk: d too large for the re-score kernel");
        if ((q->flags & B200_Q_FORCE_TC) && !use_tc)
            return fail(B200_E_UNSUPPORTED, "b200_rank_topk: tensor-core path unavailable (tc_dtype=%d, k=%d, d_pad=%d, n_pos=%lld)",
                        E->tc_dtype, k_out, E->d_pad, (long long)n_pos);
        const bool peers = shared && use_tc;  // (zero peers: the same protocol, nothing to adopt)
        if (shared && use_tc && k_out > 24) return fail(B200_E_UNSUPPORTED, "b200_rank_topk: B200_Q_SHARED_THRESHOLDS needs k <= 24");

        // failure lists (absolute rows) + counters: [fb1 | fb2 | fbA | fbB | counters]
        E->fb_rows.ensure(sizeof(int32_t) * (4 * n_rows + 16));
        int32_t* fb1 = E->fb_rows.as<int32_t>();
        int32_t* fb2 = fb1 + n_rows;
        int32_t* fbA = fb2 + n_rows;
        int32_t* fbB = fbA + n_rows;
        int32_t* cnt = fbB + n_rows;
        CK(cudaMemsetAsync(cnt, 0, 16 * sizeof(int32_t), st));
        if (wide || (use_tc && k_out > 24)) E->excl.ensure(sizeof(int32_t) * (size_t)n_rows * k_out);

        // ---------------- main pass over the rows of one chunk
        auto main_pass = [&](int64_t r0, int64_t r1) {
            const int64_t nr = r1 - r0;
            int32_t* oi = c.o_ids + r0 * k_out;
            float* os = c.o_scores + r0 * k_out;
            int32_t* oc = c.o_counts + r0;
            init_outputs_kernel<<<grid_for(std::max<int64_t>(nr * k_out, nr), 256), 256, 0, st>>>(oi, os, oc, nr, k_out);
            CK(cudaGetLastError());
            S.n_launches++;
            const float* sub = (c.sub32 && !c.rowmap) ? c.sub32 + r0 * d : c.sub32;
            const int64_t* rm = c.rowmap ? c.rowmap + r0 : nullptr;
            const int64_t* ip = c.indptr ? c.indptr + r0 : nullptr;
            if (sparse_sub) {
                S.path = 2;
                run_sparse(c, sp_indptr + r0, sp_indices, sp_data, nr, ip, oi, os, oc);
            } else if (!use_tc && k_out > 128) {
                S.path = 3;  // materialised exhaustive scores + streaming selection passes
                run_dense_large_k(c, sub, rm, ip, nr, oi, os, oc);
                if (c.o_bounds) {
                    fill_f32_kernel<<<grid_for(nr, 256), 256, 0, st>>>(c.o_bounds + r0, nr, -INFINITY);
                    CK(cudaGetLastError());
                }
            } else if (!use_tc) {
                S.path = 0;
                run_exact(c, nullptr, nr, sub, rm, ip, oi, os, oc, 0, k_out, true);
                if (c.o_bounds) {  // exhaustive lists: nothing was discarded
                    fill_f32_kernel<<<grid_for(nr, 256), 256, 0, st>>>(c.o_bounds + r0, nr, -INFINITY);
                    CK(cudaGetLastError());
                }
            } else {
                S.path = 1;
                S.tc_dtype = E->tc_dtype;
                TcPass t;
                t.n_sel = nr;
                t.sub32 = sub;
                t.rowmap = rm;
                t.indptr = ip;
                t.o_ids = oi;
                t.o_scores = os;
                t.o_counts = oc;
                t.o_bounds = c.o_bounds ? c.o_bounds + r0 : nullptr;
                t.nw = c.nw;
                t.kc = k_cand;
                t.row0 = r0;
                t.fb_list = fb1;
                t.fb_count = cnt;
                t.main = true;
                t.peers = peers;
                t.k0 = 0;
                t.kp = k_out;
                t.wide = wide;
                run_tc(c, t);
            }
            if (E->id_offset != 0) {
                add_offset_kernel<<<grid_for(nr * k_out, 256), 256, 0, st>>>(oi, nr * k_out, (int32_t)E->id_offset);
                CK(cudaGetLastError());
                S.n_launches++;
            }
        };

        // Re-rank `n_sel` rows (absolute row numbers in `rows`) without any shortcut that could fail again unnoticed:
        // k <= 24: one pass with the widest lists (32 slots, 8-warp kernel), then the exhaustive kernel for what still fails;
        // k  > 24: certified passes of 20 results with exclusion lists, each followed by its own wide-list pass and the
        // exhaustive kernel.  Results are written with LOCAL ids; the caller applies the id offset.
        auto rerank_rows = [&](const int32_t* rows, int64_t n_sel) {
            init_rows_kernel<<<grid_for(n_sel * k_out, 256), 256, 0, st>>>(c.o_ids, c.o_scores, c.o_counts, rows, n_sel, k_out);
            CK(cudaGetLastError());
            S.n_launches++;
            const int k_pass = k_out <= 24 ? k_out : 20;
            const int kc_pass = k_out <= 24 ? 32 : (c.bf16 ? 30 : 25);
            for (int k0 = 0; k0 < k_out; k0 += k_pass) {
                const int kp = std::min(k_pass, k_out - k0);
                if (k0 > 0) {
                    build_exclusion_kernel<<<grid_for(n_sel * 32, 256), 256, 0, st>>>(c.o_ids, rows, n_sel, k_out, k0, (int32_t)E->id_offset,
                                                                                     E->excl.as<int32_t>());
                    CK(cudaGetLastError());
                    S.n_launches++;
                }
                CK(cudaMemsetAsync(cnt + 2, 0, 2 * sizeof(int32_t), st));
                TcPass t;
                t.rows_dev = rows;
                t.n_sel = n_sel;
                t.sub32 = c.sub32;
                t.rowmap = c.rowmap;
                t.indptr = c.indptr;
                t.o_ids = c.o_ids;
                t.o_scores = c.o_scores;
                t.o_counts = c.o_counts;
                t.nw = 8;
                t.kc = std::min(32, std::max(kc_pass - (k_pass - kp), kp));
                t.k0 = k0;
                t.kp = kp;
                t.fb_list = fbA;
                t.fb_count = cnt + 2;
                run_tc(c, t);
                int64_t n_fb = read_counter(c, cnt + 2);
                const int32_t* f = fbA;
                if (n_fb > 0 && t.kc < 32) {  // the pass's own second chance: widest lists
                    TcPass t2 = t;
                    t2.rows_dev = fbA;
                    t2.n_sel = n_fb;
                    t2.kc = 32;
                    t2.fb_list = fbB;
                    t2.fb_count = cnt + 3;
                    run_tc(c, t2);
                    n_fb = read_counter(c, cnt + 3);
                    f = fbB;
                }
                S.n_exact_rows += n_fb;
                if (n_fb > 0) run_exact(c, f, n_fb, c.sub32, c.rowmap, c.indptr, c.o_ids, c.o_scores, c.o_counts, k0, k0 + kp, false);
            }
        };

        // ---------------- chunk pipeline
        const bool multipass_main = use_tc && k_out > 24 && !wide;
        int64_t chunk = n_rows;
        if (!in_dev && use_tc && !multipass_main) {
            const int64_t wave = (int64_t)(E->sm_count / 2) * 256;  // subject rows one wave of CTA pairs works on
            int64_t want = 8 * wave;
            if (const char* env = getenv("B200_CHUNK_ROWS")) want = std::max<int64_t>(256, atoll(env));  // test hook
            if (n_rows >= 2 * want) chunk = want;
        }
        const int64_t n_chunks = (n_rows + chunk - 1) / chunk;
        cudaStream_t cs = n_chunks > 1 ? E->cs : st;
        S.n_chunks = (int32_t)n_chunks;
        if (n_chunks > 1) {
            CK(cudaEventRecord(E->evp[2], st));  // the copy stream starts after everything queued so far (whitelist, ...)
            CK(cudaStreamWaitEvent(cs, E->evp[2], 0));
        }
        stage_rows(0, std::min(chunk, n_rows), cs);
        CK(cudaEventRecord(E->evp[0], cs));
        CK(cudaEventRecord(E->ev[1], cs));
        auto copy_back = [&](int64_t r0, int64_t r1, cudaStream_t s) {
            CK(cudaMemcpyAsync(q->out_ids + r0 * k_out, c.o_ids + r0 * k_out, sizeof(int32_t) * (r1 - r0) * k_out, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(q->out_scores + r0 * k_out, c.o_scores + r0 * k_out, sizeof(float) * (r1 - r0) * k_out, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(q->out_counts + r0, c.o_counts + r0, sizeof(int32_t) * (r1 - r0), cudaMemcpyDeviceToHost, s));
            S.d2h_bytes += (int64_t)((r1 - r0) * k_out * 8 + (r1 - r0) * 4);
            if (shared) {
                CK(cudaMemcpyAsync(q->out_bounds + r0, c.o_bounds + r0, sizeof(float) * (r1 - r0), cudaMemcpyDeviceToHost, s));
                S.d2h_bytes += (int64_t)(r1 - r0) * 4;
            }
        };
        for (int64_t ci = 0; ci < n_chunks; ++ci) {
            const int64_t r0 = ci * chunk, r1 = std::min(n_rows, r0 + chunk);
            if (ci + 1 < n_chunks) {
                stage_rows(r1, std::min(n_rows, r1 + chunk), cs);
                CK(cudaEventRecord(E->evp[(ci + 1) & 1], cs));
            }
            if (n_chunks > 1) CK(cudaStreamWaitEvent(st, E->evp[ci & 1], 0));
            if (multipass_main) {
                // every row takes the certified multi-pass route (tuning / test hook B200_WIDE=0)
                S.path = 1;
                S.tc_dtype = E->tc_dtype;
                S.k_cand = k_cand;
                S.epi_warps = 8;
                iota_kernel<<<grid_for(n_rows, 256), 256, 0, st>>>(fb1, n_rows);
                CK(cudaGetLastError());
                rerank_rows(fb1, n_rows);
                if (E->id_offset != 0) {
                    add_offset_kernel<<<grid_for(n_rows * k_out, 256), 256, 0, st>>>(c.o_ids, n_rows * k_out, (int32_t)E->id_offset);
                    CK(cudaGetLastError());
                }
            } else {
                main_pass(r0, r1);
            }
            if (ci + 1 == n_chunks) CK(cudaEventRecord(E->ev[4], st));
            if (!out_dev) {
                if (n_chunks > 1) {
                    CK(cudaEventRecord(E->evp[2], st));
                    CK(cudaStreamWaitEvent(cs, E->evp[2], 0));
                }
                copy_back(r0, r1, cs);
            }
        }
        // ---------------- rows whose certificate failed in the main pass (all chunks): re-rank, patch the results
        int64_t n_fb = 0;
        if (use_tc && !multipass_main && !peers && !sparse_sub) {
            n_fb = read_counter(c, cnt);
            S.n_fallback_rows = n_fb;
            if (n_fb > 0) {
                rerank_rows(fb1, n_fb);
                if (E->id_offset != 0) {
                    add_offset_rows_kernel<<<grid_for(n_fb * k_out, 256), 256, 0, st>>>(c.o_ids, fb1, n_fb, k_out, (int32_t)E->id_offset);
                    CK(cudaGetLastError());
                    S.n_launches++;
                }
                if (!out_dev) {
                    // packed copies of the patched rows: one more small transfer, scattered into the caller's arrays below
                    const size_t row_bytes = (size_t)k_out * 8 + 8;
                    E->patch.ensure(row_bytes * n_fb);
                    int32_t* g_ids = E->patch.as<int32_t>();
                    float* g_sc = reinterpret_cast<float*>(g_ids + n_fb * k_out);
                    int32_t* g_cnt = reinterpret_cast<int32_t*>(g_sc + n_fb * k_out);
                    int32_t* g_rows = g_cnt + n_fb;
                    gather_rows_kernel<<<grid_for(n_fb * k_out, 256), 256, 0, st>>>(c.o_ids, c.o_scores, c.o_counts, fb1, n_fb, k_out, g_ids,
                                                                                   g_sc, g_cnt);
                    CK(cudaGetLastError());
                    CK(cudaMemcpyAsync(g_rows, fb1, sizeof(int32_t) * n_fb, cudaMemcpyDeviceToDevice, st));
                    E->h_patch.resize(row_bytes * n_fb);
                    if (n_chunks > 1) {  // the caller's arrays must hold the chunk copies before they are patched
                        CK(cudaEventRecord(E->evp[2], cs));
                        CK(cudaStreamWaitEvent(st, E->evp[2], 0));
                    }
                    CK(cudaMemcpyAsync(E->h_patch.data(), E->patch.p, row_bytes * n_fb, cudaMemcpyDeviceToHost, st));
                    CK(cudaStreamSynchronize(st));
                    const int32_t* h_ids = reinterpret_cast<const int32_t*>(E->h_patch.data());
                    const float* h_sc = reinterpret_cast<const float*>(h_ids + n_fb * k_out);
                    const int32_t* h_cnt = reinterpret_cast<const int32_t*>(h_sc + n_fb * k_out);
                    const int32_t* h_rows = h_cnt + n_fb;
                    for (int64_t i = 0; i < n_fb; ++i) {
                        const int64_t r = h_rows[i];
                        memcpy(q->out_ids + r * k_out, h_ids + i * k_out, sizeof(int32_t) * k_out);
                        memcpy(q->out_scores + r * k_out, h_sc + i * k_out, sizeof(float) * k_out);
                        q->out_counts[r] = h_cnt[i];
                    }
                    S.d2h_bytes += (int64_t)(row_bytes * n_fb);
                }
            }
        }
        if (n_chunks > 1) {  // the main stream (and through it the caller) sees the copies of the last chunks
            CK(cudaEventRecord(E->evp[2], cs));
            CK(cudaStreamWaitEvent(st, E->evp[2], 0));
        }
        CK(cudaEventRecord(E->ev[5], st));
        if (user && out_dev) {
            CK(cudaEventRecord(E->ev[7], st));
            CK(cudaStreamWaitEvent(user, E->ev[7], 0));
        }
        CK(cudaStreamSynchronize(st));
        CK(cudaEventElapsedTime(&S.ms_total, E->ev[0], E->ev[5]));
        CK(cudaEventElapsedTime(&S.ms_h2d, E->ev[0], E->ev[1]));  // exposed part: the first chunk's inputs
        CK(cudaEventElapsedTime(&S.ms_d2h, E->ev[4], E->ev[5]));  // exposed part: the last chunk's results (+ re-ranked rows)
        c.collect_times();
    } catch (const CudaError& ce) {
        return fail(ce.e == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA, "b200_rank_topk: %s failed at line %d: %s", ce.what,
                    ce.line, cudaGetErrorString(ce.e));
    }
    if (stats) *stats = c.S;
    return B200_OK;
}

static int merge_impl(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids, const float* scores,
                      const int32_t* counts, const float* bounds, int64_t list_stride, int32_t* out_ids, float* out_scores,
                      int32_t* out_counts, int32_t* fail_rows, int32_t* fail_count) {
    if (n_lists <= 0 || n_rows < 0 || k <= 0 || !ids || !scores || !counts || !out_ids || !out_scores || !out_counts)
        return fail(B200_E_INVALID, "b200_rank_merge: bad arguments");
    if (bounds && (!fail_rows || !fail_count)) return fail(B200_E_INVALID, "b200_rank_merge_certified: fail_rows / fail_count are NULL");
    if (bounds && k > 32) return fail(B200_E_UNSUPPORTED, "b200_rank_merge_certified: k <= 32");
    if (n_rows == 0) return B200_OK;
    try {
        CK(cudaSetDevice(device));
        cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
        init_outputs_kernel<<<grid_for(n_rows * k, 256), 256, 0, st>>>(out_ids, out_scores, out_counts, n_rows, k);
        CK(cudaGetLastError());
        for (int k0 = 0; k0 < k; k0 += 32) {
            SelectParams sp{};
            sp.in_scores = scores;
            sp.in_ids = ids;
            sp.in_counts = counts;
            sp.in_bounds = bounds;
            sp.n_lists = n_lists;
            sp.L = k;
            sp.n_sel = n_rows;
            sp.list_stride_rows = n_rows;
            sp.list_stride_elems = list_stride;
            sp.k_out = k;
            sp.k0 = k0;
            sp.kp = std::min(32, k - k0);
            sp.out_ids = out_ids;
            sp.out_scores = out_scores;
            sp.out_counts = out_counts;
            sp.fb_rows = fail_rows;
            sp.fb_count = fail_count;
            merge_select_kernel<<<grid_for(n_rows, SEL_WARPS), SEL_WARPS * 32, 0, st>>>(sp);
            CK(cudaGetLastError());
        }
    } catch (const CudaError& ce) {
        return fail(B200_E_CUDA, "b200_rank_merge: %s failed: %s", ce.what, cudaGetErrorString(ce.e));
    }
    return B200_OK;
}

int b200_rank_merge(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids,
                    const float* scores, const int32_t* counts, int32_t* out_ids, float* out_scores, int32_t* out_counts) {
    return merge_impl(device, stream, n_lists, n_rows, k, ids, scores, counts, nullptr, 0, out_ids, out_scores, out_counts, nullptr, nullptr);
}

int b200_rank_merge_certified(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids,
                              const float* scores, const int32_t* counts, const float* bounds, int64_t list_stride, int32_t* out_ids,
                              float* out_scores, int32_t* out_counts, int32_t* fail_rows, int32_t* fail_count) {
    if (!bounds) return fail(B200_E_INVALID, "b200_rank_merge_certified: bounds is NULL");
    return merge_impl(device, stream, n_lists, n_rows, k, ids, scores, counts, bounds, list_stride, out_ids, out_scores, out_counts,
                      fail_rows, fail_count);
}

}  // extern "C"
