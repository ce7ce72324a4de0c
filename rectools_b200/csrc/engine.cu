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
#include "tc_topk.cuh"
#include "tc2_topk.cuh"
#include "tc3_topk.cuh"
#include "tc4_topk.cuh"

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
bool make_tensor_map(CUtensorMap* tm, const void* base, int64_t rows, int d_pad, bool is_bf16, int box_rows = 128) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return false;
    cuuint64_t dims[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)d_pad * 2};
    cuuint32_t box[2] = {(cuuint32_t)b200::tc::KBLK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

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

    // resident object data
    DevBuf obj32;     // [n_obj, d] fp32 master copy
    DevBuf obj16;     // [n_obj_pad, d_pad] fp16 / bf16, pre-scaled (and pre-normalised for COSINE)
    DevBuf obj_norms; // [n_obj] fp32 (COSINE)
    int obj_exp = 0;
    int64_t id_offset = 0;
    float max_obj_norm = 0.f;
    bool obj32_owned = true;
    const float* obj32_ptr = nullptr;
    CUtensorMap tm_obj_full;
    bool tm_obj_ok = false;
    bool tc4_ok = false;  // experimental tc4_topk_kernel usable (B200_TC_KERNEL=4)

    // resident subjects (optional)
    DevBuf sub32_res;
    int64_t n_sub_res = 0;
    const float* sub32_res_ptr = nullptr;

    // per-call staging / workspace
    DevBuf sub32, sub16, row_exp, rowmap, indptr, indices, wl, obj16_wl;
    DevBuf out_ids, out_scores, out_counts;
    DevBuf cand_scores, cand_ids, cand_counts;
    DevBuf part_scores, part_ids;
    DevBuf fb_rows, scratch, excl, carousel;
    int32_t* h_pinned = nullptr;  // small pinned scratch (fallback count)

    size_t hbm_bytes() const {
        const DevBuf* all[] = {&obj32, &obj16, &obj_norms, &sub32_res, &sub32, &sub16, &row_exp, &rowmap, &indptr,
                               &indices, &wl, &obj16_wl, &out_ids, &out_scores, &out_counts, &cand_scores, &cand_ids,
                               &cand_counts, &part_scores, &part_ids, &fb_rows, &scratch, &excl, &carousel};
        size_t t = 0;
        for (auto* b : all) t += b->cap;
        return t;
    }
    void free_all() {
        DevBuf* all[] = {&obj32, &obj16, &obj_norms, &sub32_res, &sub32, &sub16, &row_exp, &rowmap, &indptr,
                         &indices, &wl, &obj16_wl, &out_ids, &out_scores, &out_counts, &cand_scores, &cand_ids,
                         &cand_counts, &part_scores, &part_ids, &fb_rows, &scratch, &excl, &carousel};
        for (auto* b : all) b->release();
        if (h_pinned) cudaFreeHost(h_pinned);
        h_pinned = nullptr;
        for (auto& e : ev)
            if (e) cudaEventDestroy(e);
        for (auto& e : evp)
            if (e) cudaEventDestroy(e);
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

    if (tc_mode == B200_TC_OFF || E->cc_major != 10) {
        E->tc_dtype = B200_TC_OFF;
        return;
    }
    E->tc_dtype = (tc_mode == B200_TC_BF16) ? B200_TC_BF16 : B200_TC_FP16;
    E->obj_exp = (E->tc_dtype == B200_TC_FP16) ? fp16_scale_exp(absmax) : 0;
    E->n_obj_pad = round_up(std::max<int64_t>(n, 1), tc::TILE_N);
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
    E->tm_obj_ok = make_tensor_map(&E->tm_obj_full, E->obj16.p, E->n_obj_pad, E->d_pad, E->tc_dtype == B200_TC_BF16);
    if (!E->tm_obj_ok) E->tc_dtype = B200_TC_OFF;
}

struct TcPlan {
    int s_sub, kblocks, n_stages, smem_bytes;
    bool ok;
};

TcPlan plan_tc(int d_pad) {
    TcPlan pl{};
    pl.kblocks = d_pad / tc::KBLK;
    for (int s = 2; s >= 1; --s) {
        const int a = s * pl.kblocks * tc::BLK_BYTES;
        const int lists = s * tc::TILE_M * 32 * 8;
        const int fixed = a + lists + 1024 /*alignment slack*/ + 512 /*barriers*/;
        const int avail = tc::SMEM_LIMIT - fixed;
        int stages = avail / tc::BLK_BYTES;
        if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
        if (stages >= 3 || (s == 1 && stages >= 2)) {
            pl.s_sub = s;
            pl.n_stages = stages;
            pl.smem_bytes = fixed + stages * tc::BLK_BYTES;
            pl.ok = true;
            return pl;
        }
    }
    pl.ok = false;
    return pl;
}

// Shared-memory plan of the 2-SM kernel (per CTA: 128 subject rows, half of every object tile).
TcPlan plan_tc2(int d_pad, int tile_n) {
    TcPlan pl{};
    pl.kblocks = d_pad / tc::KBLK;
    pl.s_sub = 2;  // two candidate lists per row (one per column half)
    const int a = pl.kblocks * tc::BLK_BYTES;
    const int lists = 2 * tc::TILE_M * 32 * 8 + 2 * tc::TILE_M * 8;
    const int fixed = a + lists + 1024 /*alignment slack*/ + 512 /*barriers*/;
    const int blkb = tile_n / 2 * tc::KBLK * 2;  // object block bytes per CTA
    int stages = (tc::SMEM_LIMIT - fixed) / blkb;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    pl.ok = stages >= 2;
    pl.n_stages = stages;
    pl.smem_bytes = fixed + stages * blkb;
    return pl;
}

// Shared-memory plan of the default kernel (tc3_topk.cuh): as plan_tc2(d_pad, 256) plus the deferred-hit FIFOs.
TcPlan plan_tc3(int d_pad) {
    TcPlan pl{};
    pl.kblocks = d_pad / tc::KBLK;
    pl.s_sub = 2;
    const int a = pl.kblocks * tc::BLK_BYTES;
    const int lists = 2 * tc::TILE_M * 32 * 8 + 2 * tc::TILE_M * 8 + tc::T3_QBYTES;
    const int fixed = a + lists + 1024 /*alignment slack*/ + 512 /*barriers*/;
    int stages = (tc::SMEM_LIMIT - fixed) / tc::BLK_BYTES;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    pl.ok = stages >= 2;
    pl.n_stages = stages;
    pl.smem_bytes = fixed + stages * tc::BLK_BYTES;
    return pl;
}

// Shared-memory plan of the experimental 16-epilogue-warp kernel (tc4_topk.cuh).
TcPlan plan_tc4(int d_pad) {
    TcPlan pl{};
    pl.kblocks = d_pad / tc::KBLK;
    pl.s_sub = 2;
    const int a = pl.kblocks * tc::BLK_BYTES;
    const int lists = 2 * tc::T4_LIST_BYTES + 4 * tc::TILE_M * 8 + tc::T4_QBYTES;
    const int fixed = a + lists + 1024 /*alignment slack*/ + 512 /*barriers*/;
    int stages = (tc::SMEM_LIMIT - fixed) / tc::BLK_BYTES;
    if (stages > tc::MAX_STAGES) stages = tc::MAX_STAGES;
    pl.ok = stages >= 2;
    pl.n_stages = stages;
    pl.smem_bytes = fixed + stages * tc::BLK_BYTES;
    return pl;
}

uint32_t make_idesc2(bool bf16, int tile_n) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (bf16 ? 1u : 0u) << 7;
    d |= (bf16 ? 1u : 0u) << 10;
    d |= (uint32_t)(tile_n >> 3) << 17;
    d |= (uint32_t)(256 >> 4) << 24;          // M = 256 across the CTA pair
    return d;
}

uint32_t make_idesc(bool bf16) {
    uint32_t d = 0;
    d |= 1u << 4;                       // accumulator format: F32
    d |= (bf16 ? 1u : 0u) << 7;         // A format
    d |= (bf16 ? 1u : 0u) << 10;        // B format
    // bits 13/14: no negate; bits 15/16: both operands K-major
    d |= (uint32_t)(tc::TILE_N >> 3) << 17;
    d |= (uint32_t)(tc::TILE_M >> 4) << 24;
    return d;
}

}  // namespace

extern "C" {

const char* b200_rank_last_error(void) { return g_last_error.c_str(); }
int b200_rank_abi_version(void) { return B200_RANK_ABI_VERSION; }

int b200_rank_create(b200_rank_engine** out, const float* objects, int64_t n_objects, int32_t d, int32_t distance,
                     int32_t device, int32_t tc_mode, int32_t flags) {
    if (!out) return fail(B200_E_INVALID, "b200_rank_create: out is NULL");
    *out = nullptr;
    if (n_objects < 0 || d <= 0 || (!objects && n_objects > 0))
        return fail(B200_E_INVALID, "b200_rank_create: bad object matrix (n=%lld, d=%d)", (long long)n_objects, d);
    if (n_objects >= (1ll << 31) - 1) return fail(B200_E_UNSUPPORTED, "b200_rank_create: more than 2^31-2 objects");
    if (distance != B200_DIST_DOT && distance != B200_DIST_COSINE)
        return fail(B200_E_INVALID, "b200_rank_create: distance must be B200_DIST_DOT or B200_DIST_COSINE");
    if (tc_mode < B200_TC_AUTO || tc_mode > B200_TC_OFF) return fail(B200_E_INVALID, "b200_rank_create: bad tc_mode");
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
        CK(cudaMallocHost(&E->h_pinned, 64));
        if (flags & B200_F_OBJECTS_ON_DEVICE) {
            E->obj32_ptr = objects;
            E->obj32_owned = false;
        } else {
            E->obj32.ensure(sizeof(float) * std::max<int64_t>(n_objects * d, 1));
            if (n_objects > 0)
                CK(cudaMemcpyAsync(E->obj32.p, objects, sizeof(float) * n_objects * d, cudaMemcpyHostToDevice, E->st));
            E->obj32_ptr = E->obj32.as<float>();
        }
        if (tc_mode != B200_TC_OFF && E->d_pad > 1024) tc_mode = B200_TC_OFF;
        prepare_objects(E, tc_mode);
        if (E->tc_dtype != B200_TC_OFF) {
            TcPlan pl = plan_tc(E->d_pad);
            if (!pl.ok) {
                E->tc_dtype = B200_TC_OFF;
            } else {
                CK(cudaFuncSetAttribute(tc::tc_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl.smem_bytes));
            }
            TcPlan pl2 = plan_tc2(E->d_pad, 256), pl2b = plan_tc2(E->d_pad, 128), pl3 = plan_tc3(E->d_pad);
            if (pl3.ok) CK(cudaFuncSetAttribute(tc::tc3_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl3.smem_bytes));
            TcPlan pl4 = plan_tc4(E->d_pad);  // experimental kernel: never allowed to fail the engine
            if (pl4.ok) {
                if (cudaFuncSetAttribute(tc::tc4_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pl4.smem_bytes) == cudaSuccess)
                    E->tc4_ok = true;
                else
                    (void)cudaGetLastError();
            }
            if (pl2.ok) {
                CK(cudaFuncSetAttribute(tc::tc2_topk_kernel<256, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl2.smem_bytes));
                CK(cudaFuncSetAttribute(tc::tc2_topk_kernel<256, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl2.smem_bytes));
            }
            if (pl2b.ok) {
                CK(cudaFuncSetAttribute(tc::tc2_topk_kernel<128, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl2b.smem_bytes));
                CK(cudaFuncSetAttribute(tc::tc2_topk_kernel<128, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, pl2b.smem_bytes));
            }
        }
        CK(cudaFuncSetAttribute(select_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
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

int b200_rank_topk(b200_rank_engine* E, const b200_rank_query* q, b200_rank_stats* stats) {
    if (!E || !q) return fail(B200_E_INVALID, "b200_rank_topk: NULL argument");
    if (q->n_rows < 0) return fail(B200_E_INVALID, "b200_rank_topk: n_rows < 0");
    if (q->k <= 0) return fail(B200_E_INVALID, "b200_rank_topk: k must be positive");
    if (!q->subjects && !q->subject_ids) return fail(B200_E_INVALID, "b200_rank_topk: neither subjects nor subject_ids given");
    if (!q->subjects && !E->sub32_res_ptr)
        return fail(B200_E_INVALID, "b200_rank_topk: subject_ids given but b200_rank_set_subjects was never called");
    if (q->subjects && q->subject_ids && q->n_subjects_total <= 0)
        return fail(B200_E_INVALID, "b200_rank_topk: subjects + subject_ids need n_subjects_total");
    if (q->whitelist && q->n_whitelist < 0) return fail(B200_E_INVALID, "b200_rank_topk: n_whitelist < 0");
    if (q->n_rows > 0 && (!q->out_ids || !q->out_scores || !q->out_counts))
        return fail(B200_E_INVALID, "b200_rank_topk: output pointers are NULL");
    if (q->n_rows >= (1ll << 31) - 64) return fail(B200_E_UNSUPPORTED, "b200_rank_topk: more than 2^31 rows per call");
    if ((q->flags & B200_Q_FORCE_EXACT) && (q->flags & B200_Q_FORCE_TC))
        return fail(B200_E_INVALID, "b200_rank_topk: FORCE_EXACT and FORCE_TC are exclusive");

    std::lock_guard<std::mutex> lock(E->mu);
    b200_rank_stats S;
    memset(&S, 0, sizeof(S));
    const int64_t n_rows = q->n_rows;
    const int64_t n_pos = q->whitelist ? q->n_whitelist : E->n_obj;
    const int k_out = (int)std::min<int64_t>(q->k, n_pos);
    S.k_out = k_out;
    const bool in_dev = q->flags & B200_Q_INPUTS_ON_DEVICE;
    const bool out_dev = q->flags & B200_Q_OUTPUTS_ON_DEVICE;
    const int d = E->d;
    if (n_rows == 0 || k_out <= 0) {
        if (stats) *stats = S;
        return B200_OK;
    }
    try {
        CK(cudaSetDevice(E->device));
        cudaStream_t st = E->st;
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
        const float* sub32 = nullptr;
        const int64_t* rowmap = nullptr;
        auto stage = [&](DevBuf& buf, const void* src, size_t bytes) -> const void* {  // un-chunked items, main stream
            if (in_dev) return src;
            buf.ensure(std::max<size_t>(bytes, 16));
            if (bytes) CK(cudaMemcpyAsync(buf.p, src, bytes, cudaMemcpyHostToDevice, st));
            S.h2d_bytes += (int64_t)bytes;
            return buf.p;
        };
        const bool chunk_subjects = q->subjects && !q->subject_ids && !in_dev;  // subject rows arrive in batch order
        if (q->subjects) {
            const int64_t rows_in = q->subject_ids ? q->n_subjects_total : n_rows;
            if (chunk_subjects) {
                E->sub32.ensure(std::max<size_t>(sizeof(float) * rows_in * d, 16));
                sub32 = E->sub32.as<float>();
            } else {
                sub32 = (const float*)stage(E->sub32, q->subjects, sizeof(float) * rows_in * d);
            }
        } else {
            sub32 = E->sub32_res_ptr;
        }
        if (q->subject_ids) {
            if (in_dev) {
                rowmap = q->subject_ids;
            } else {
                E->rowmap.ensure(std::max<size_t>(sizeof(int64_t) * n_rows, 16));
                rowmap = E->rowmap.as<int64_t>();
            }
        }
        const int64_t* indptr = nullptr;
        const int32_t* indices = nullptr;
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
                indptr = q->csr_indptr;
                indices = q->csr_indices;
            } else {
                E->indptr.ensure(sizeof(int64_t) * (n_rows + 1));
                E->indices.ensure(std::max<size_t>(sizeof(int32_t) * nnz, 16));
                indptr = E->indptr.as<int64_t>();
                indices = E->indices.as<int32_t>();
            }
            if (nnz == 0) indptr = nullptr;  // an all-empty filter is no filter (cf. rank_implicit.py:169-173)
        }
        const int32_t* wl = nullptr;
        if (q->whitelist) wl = (const int32_t*)stage(E->wl, q->whitelist, sizeof(int32_t) * n_pos);
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
            if (indptr) {
                h2d(E->indptr.as<int64_t>() + r0, q->csr_indptr + r0, sizeof(int64_t) * (r1 - r0 + 1));
                const int64_t z0 = q->csr_indptr[r0], z1 = q->csr_indptr[r1];
                if (z1 < z0) throw CudaError{cudaErrorInvalidValue, "csr_indptr must be non-decreasing", __LINE__};
                h2d(E->indices.as<int32_t>() + z0, q->csr_indices + z0, sizeof(int32_t) * (z1 - z0));
            }
            S.h2d_bytes += (int64_t)bytes;
        };

        // ---------------- outputs
        int32_t* o_ids;
        float* o_scores;
        int32_t* o_counts;
        if (out_dev) {
            o_ids = q->out_ids;
            o_scores = q->out_scores;
            o_counts = q->out_counts;
        } else {
            E->out_ids.ensure(sizeof(int32_t) * n_rows * k_out);
            E->out_scores.ensure(sizeof(float) * n_rows * k_out);
            E->out_counts.ensure(sizeof(int32_t) * n_rows);
            o_ids = E->out_ids.as<int32_t>();
            o_scores = E->out_scores.as<float>();
            o_counts = E->out_counts.as<int32_t>();
        }
        // ---------------- path choice
        TcPlan pl = plan_tc(E->d_pad);
        // 2-SM kernel (CTA pairs, cta_group::2) unless disabled or impossible; B200_TC_KERNEL=1 selects the 1-SM kernel
        // B200_TC_KERNEL: 1 = 1-SM kernel (tc_topk.cuh), 2 = previous 2-SM kernel (tc2_topk.cuh), default 3 = tc3_topk.cuh,
        // 4 = experimental tc4_topk.cuh for passes with K' <= 16 (everything else of such a call runs on tc3)
        bool use_2sm = (E->sm_count % 2 == 0);
        int kernel_gen = 3;
        if (const char* env = getenv("B200_TC_KERNEL")) {
            kernel_gen = atoi(env);
            use_2sm = use_2sm && kernel_gen != 1;
        }
        // tile width of the 2-SM kernel: 256 objects x 2 TMEM buffers (default) or 128 x 4 (B200_TC_TILE=128)
        int tile2_n = 256;
        if (const char* env = getenv("B200_TC_TILE")) tile2_n = atoi(env) == 128 ? 128 : 256;
        bool use_gen3 = use_2sm && kernel_gen != 2 && tile2_n == 256;
        if (use_gen3) {
            TcPlan pl3 = plan_tc3(E->d_pad);
            if (pl3.ok)
                pl = pl3;
            else
                use_gen3 = false;
        }
        const TcPlan pl4 = plan_tc4(E->d_pad);
        const bool use_gen4 = use_gen3 && kernel_gen == 4 && pl4.ok && E->tc4_ok;
        if (use_2sm && !use_gen3) {
            TcPlan pl2 = plan_tc2(E->d_pad, tile2_n);
            if (pl2.ok)
                pl = pl2;
            else
                use_2sm = false;
        }
        // Candidates kept per list by the tensor-core pass (K' >= k; the surplus is the certificate's safety margin).
        // 1-SM kernel: one list per row.  2-SM kernel: two lists per row (one per column half), so a small surplus per
        // list already gives ~2k candidates; rows where (nearly) all of the top-k fall into one half fail the
        // certificate and take the second-chance pass.  Inserts, the dominant epilogue cost, scale with K'.
        int k_cand = 0;
        const bool bf16_tc = E->tc_dtype == B200_TC_BF16;
        if (use_gen4 && k_out <= 24) {
            // four lists per row: a list may be SHORTER than k (the certificate only needs the k-th exact score above every
            // full list's minimum); rows whose top-k crowd into one column quarter take the second-chance pass
            k_cand = std::min(tc::T4_SLOTS, (k_out <= 10 ? 8 : k_out <= 16 ? 12 : 16) + (bf16_tc ? 2 : 0));
        } else if (use_2sm) {
            const int surplus = bf16_tc ? std::max(6, k_out / 2) : std::max(2, k_out / 4);
            if (k_out <= 24) k_cand = std::min(32, k_out + surplus);
            else if (k_out <= 128) k_cand = 25;  // multi-pass, see below
        } else {
            if (k_out <= 10 && !bf16_tc)
                k_cand = 16;
            else if (k_out <= 128)
                k_cand = 32;
        }
        if (const char* env = getenv("B200_TC_KCAND")) {  // tuning hook
            const int forced = atoi(env);
            if ((forced >= k_out || use_gen4) && forced >= 4 && forced <= 32) k_cand = forced;
        }
        bool use_tc = E->tc_dtype != B200_TC_OFF && pl.ok && k_cand > 0 && !(q->flags & B200_Q_FORCE_EXACT) &&
                      n_pos >= (int64_t)k_cand * 4;
        if (use_tc && !(q->flags & B200_Q_FORCE_TC)) {
            // tiny problems are cheaper (and exercised) on the exhaustive kernel
            if ((double)n_rows * (double)n_pos < 4.0e6) use_tc = false;
        }
        if (use_tc && (size_t)SEL_WARPS * d * sizeof(float) > 64 * 1024)
            return fail(B200_E_UNSUPPORTED, "b200_rank_topk: d too large for the re-score kernel");
        if ((q->flags & B200_Q_FORCE_TC) && !use_tc)
            return fail(B200_E_UNSUPPORTED, "b200_rank_topk: tensor-core path unavailable (tc_dtype=%d, k=%d, d_pad=%d, n_pos=%lld)",
                        E->tc_dtype, k_out, E->d_pad, (long long)n_pos);

        // ---------------- rank the rows of one chunk (the parameters shadow the whole-call values of the same names)
        bool wl_gathered = false;  // the whitelist gather of the 16-bit objects is shared by all chunks / passes of a call
        auto compute_rows = [&](int64_t n_rows, const float* sub32, const int64_t* rowmap, const int64_t* indptr, int32_t* o_ids,
                                float* o_scores, int32_t* o_counts) {
            init_outputs_kernel<<<grid_for(std::max<int64_t>(n_rows * k_out, n_rows), 256), 256, 0, st>>>(o_ids, o_scores, o_counts,
                                                                                                       n_rows, k_out);
            CK(cudaGetLastError());
            S.n_launches++;

            const bool cosine = E->distance == B200_DIST_COSINE;
            const float* norms = cosine ? E->obj_norms.as<float>() : nullptr;

            // exhaustive fp64 passes over `n_sel` rows (rows_dev == nullptr: all rows)
            auto run_exact = [&](const int32_t* rows_dev, int64_t n_sel, bool timed, int k_begin, int k_end) {
                const int64_t tiles_total = (n_pos + 31) / 32;
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
                    p.pos2obj = wl;
                    p.n_pos = n_pos;
                    p.d = d;
                    p.obj_norms = norms;
                    p.indptr = indptr;
                    p.indices = indices;
                    p.id_off = (int32_t)E->id_offset;
                    p.k_out = k_out;
                    p.k0 = k0;
                    p.kp = kp;
                    p.out_ids = o_ids;
                    p.out_scores = o_scores;
                    p.out_counts = o_counts;
                    p.part_scores = E->part_scores.as<float>();
                    p.part_ids = E->part_ids.as<int32_t>();
                    p.part_stride_rows = n_sel;
                    if (timed && k0 == k_begin) CK(cudaEventRecord(E->ev[2], st));
                    exact_topk_kernel<<<dim3(blocks_x, n_splits), EX_THREADS, 0, st>>>(p);
                    CK(cudaGetLastError());
                    if (timed && k0 == k_begin) CK(cudaEventRecord(E->ev[3], st));
                    SelectParams sp{};
                    sp.in_scores = E->part_scores.as<float>();
                    sp.in_ids = E->part_ids.as<int32_t>();
                    sp.in_counts = nullptr;
                    sp.n_lists = n_splits;
                    sp.L = LIST_LEN;
                    sp.n_sel = n_sel;
                    sp.list_stride_rows = n_sel;
                    sp.rows = rows_dev;
                    sp.k_out = k_out;
                    sp.k0 = k0;
                    sp.kp = kp;
                    sp.out_ids = o_ids;
                    sp.out_scores = o_scores;
                    sp.out_counts = o_counts;
                    select_kernel<false><<<grid_for(n_sel, SEL_WARPS), SEL_WARPS * 32, 0, st>>>(sp);
                    CK(cudaGetLastError());
                    S.n_launches += 2;
                }
                S.n_splits = n_splits;
            };

            // One tensor-core candidate pass + fp64 re-score + certificate over `n_sel` rows (rows_dev == nullptr: all rows).
            // Rows whose certificate fails are appended to `fb_list`; returns their number.
            auto run_tc = [&](const int32_t* rows_dev, int64_t n_sel, int kc, int k0, int kp, int32_t* fb_list, int32_t* fb_count,
                              bool timed) -> int64_t {
                const bool bf16 = E->tc_dtype == B200_TC_BF16;
                const bool g4 = use_gen4 && kc <= tc::T4_SLOTS;  // this pass on the experimental kernel
                const TcPlan& plx = g4 ? pl4 : pl;
                const int rows_per_cta = plx.s_sub * tc::TILE_M;
                const int64_t rows_pad = round_up(n_sel, rows_per_cta);
                // subjects -> 16-bit, per-row power-of-two scale
                E->sub16.ensure((size_t)rows_pad * E->d_pad * 2);
                E->row_exp.ensure(sizeof(int32_t) * rows_pad);
                {
                    const int grid = grid_for(rows_pad * 32, 256);
                    if (!bf16)
                        convert_rows_kernel<__half, true><<<grid, 256, 0, st>>>(sub32, rowmap, rows_dev, n_sel, rows_pad, d, E->d_pad, nullptr,
                                                                                0, 1, E->sub16.as<__half>(), E->row_exp.as<int32_t>());
                    else
                        convert_rows_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(sub32, rowmap, rows_dev, n_sel, rows_pad, d, E->d_pad,
                                                                                       nullptr, 0, 0, E->sub16.as<__nv_bfloat16>(),
                                                                                       E->row_exp.as<int32_t>());
                    CK(cudaGetLastError());
                    S.n_launches++;
                }
                // objects: resident 16-bit copy, or a whitelist gather of it
                const int obj_box_rows = use_2sm ? tile2_n / 2 : 128;  // object rows one CTA loads per ring block
                const void* obj_base = E->obj16.p;
                int64_t obj_rows = E->n_obj_pad;
                if (wl) {
                    const int64_t npad = round_up(n_pos, tc::TILE_N);
                    if (!wl_gathered) {  // shared by every chunk / pass / re-rank of the call
                        wl_gathered = true;
                        E->obj16_wl.ensure((size_t)npad * E->d_pad * 2);
                        const int chunks = E->d_pad * 2 / 16;
                        gather_rows16_kernel<<<grid_for(npad * chunks, 256), 256, 0, st>>>(E->obj16.as<uint4>(), wl, n_pos, npad, chunks,
                                                                                         E->obj16_wl.as<uint4>());
                        CK(cudaGetLastError());
                        S.n_launches++;
                    }
                    obj_base = E->obj16_wl.p;
                    obj_rows = npad;
                }
                CUtensorMap tm_obj, tm_sub;
                if (!make_tensor_map(&tm_obj, obj_base, obj_rows, E->d_pad, bf16, obj_box_rows) ||
                    !make_tensor_map(&tm_sub, E->sub16.p, rows_pad, E->d_pad, bf16))
                    throw CudaError{cudaErrorUnknown, "cuTensorMapEncodeTiled", __LINE__};

                tc::TcParams tp{};
                tp.s_sub = plx.s_sub;
                tp.kblocks = plx.kblocks;
                tp.n_stages = plx.n_stages;
                tp.k_cand = kc;
                tp.n_rows = n_sel;
                tp.n_pos = n_pos;
                tp.n_row_tiles = (int)(rows_pad / rows_per_cta);
                const int tile_n = use_2sm ? tile2_n : tc::TILE_N;
                const int lists_per_split = g4 ? 4 : use_2sm ? 2 : 1;
                tp.n_obj_tiles = (int)((n_pos + tile_n - 1) / tile_n);
                // object splits: fill the machine when there are few row tiles, even out the last wave otherwise
                int best_splits = 1;
                double best_eff = -1.0;
                const int max_splits = std::max(1, std::min(16, tp.n_obj_tiles * (tile_n / 128) / 32));
                const int n_units = use_2sm ? E->sm_count / 2 : E->sm_count;  // CTAs or CTA pairs working concurrently
                for (int s = 1; s <= max_splits; ++s) {
                    const double work = (double)tp.n_row_tiles * s;
                    const double waves = std::ceil(work / n_units);
                    const double eff = work / (waves * n_units) - 0.01 * (s - 1);
                    if (eff > best_eff + 1e-9) {
                        best_eff = eff;
                        best_splits = s;
                    }
                }
                if (const char* env = getenv("B200_TC_SPLITS")) {  // tuning / test hook
                    const int forced = atoi(env);
                    if (forced >= 1 && forced <= max_splits) best_splits = forced;
                }
                tp.n_splits = best_splits;
                tp.tiles_per_split = (tp.n_obj_tiles + best_splits - 1) / best_splits;
                tp.idesc = use_2sm ? make_idesc2(bf16, tile2_n) : make_idesc(bf16);
                tp.pos2obj = wl;
                tp.indptr = indptr;
                tp.indices = indices;
                tp.row_ids = rows_dev;
                if (k0 > 0) {  // objects returned by earlier passes are excluded like viewed ones
                    tp.excl = E->excl.as<int32_t>();
                    tp.excl_stride = k_out;
                    tp.excl_n = k0;
                }
                tp.id_off = (int32_t)E->id_offset;
                const int n_lists = best_splits * lists_per_split;
                E->cand_scores.ensure(sizeof(float) * (size_t)n_lists * rows_pad * 32);
                E->cand_ids.ensure(sizeof(int32_t) * (size_t)n_lists * rows_pad * 32);
                E->cand_counts.ensure(sizeof(int32_t) * (size_t)n_lists * rows_pad);
                tp.cand_scores = E->cand_scores.as<float>();
                tp.cand_ids = E->cand_ids.as<int32_t>();
                tp.cand_counts = E->cand_counts.as<int32_t>();
                tp.rows_pad = rows_pad;
                if (const char* env = getenv("B200_TC_DEBUG")) tp.debug_mode = atoi(env);  // measurement hook, results are invalid
                if (timed) S.n_splits = best_splits;
                const int n_work = tp.n_row_tiles * tp.n_splits;
                bool carousel = use_2sm;  // B200_TC_CAROUSEL=0 disables it (every work item then starts at its first object tile)
                if (const char* env = getenv("B200_TC_CAROUSEL")) carousel = carousel && atoi(env) != 0;
                if (carousel) {
                    const int n_pairs_run = std::min(n_work, n_units);
                    const int per_pair = (n_work + n_pairs_run - 1) / n_pairs_run;
                    const size_t n_ints = (size_t)best_splits + (size_t)n_pairs_run * per_pair;
                    E->carousel.ensure(sizeof(int32_t) * n_ints);
                    std::vector<int32_t> init(best_splits);
                    for (int sidx = 0; sidx < best_splits; ++sidx) init[sidx] = sidx * tp.tiles_per_split;
                    CK(cudaMemsetAsync(E->carousel.p, 0xFF, sizeof(int32_t) * n_ints, st));
                    CK(cudaMemcpyAsync(E->carousel.p, init.data(), sizeof(int32_t) * best_splits, cudaMemcpyHostToDevice, st));
                    CK(cudaStreamSynchronize(st));  // `init` is a stack buffer
                    tp.front = E->carousel.as<int32_t>();
                    tp.starts = tp.front + best_splits;
                    tp.starts_stride = per_pair;
                }
                if (timed) CK(cudaEventRecord(E->ev[2], st));
                if (use_2sm) {
                    const int grid = 2 * std::min(n_work, n_units);
                    bool stage_regs = true;  // B200_TC_STAGE=0: scan straight from TMEM in 32-column chunks
                    if (const char* env = getenv("B200_TC_STAGE")) stage_regs = atoi(env) != 0;
                    if (g4)
                        tc::tc4_topk_kernel<<<grid, tc::T4_THREADS, plx.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                    else if (use_gen3)
                        tc::tc3_topk_kernel<<<grid, tc::T3_THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                    else if (tile2_n == 256 && stage_regs)
                        tc::tc2_topk_kernel<256, 2, true><<<grid, tc::Tc2Threads<true>::THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                    else if (tile2_n == 256)
                        tc::tc2_topk_kernel<256, 2, false><<<grid, tc::Tc2Threads<false>::THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                    else if (stage_regs)
                        tc::tc2_topk_kernel<128, 4, true><<<grid, tc::Tc2Threads<true>::THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                    else
                        tc::tc2_topk_kernel<128, 4, false><<<grid, tc::Tc2Threads<false>::THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                } else {
                    const int grid = std::min(n_work, n_units);
                    tc::tc_topk_kernel<<<grid, tc::NUM_THREADS, pl.smem_bytes, st>>>(tm_sub, tm_obj, tp);
                }
                CK(cudaGetLastError());
                if (timed) CK(cudaEventRecord(E->ev[3], st));
                S.n_launches++;

                // fp64 re-score of the candidates + certificate
                CK(cudaMemsetAsync(fb_count, 0, sizeof(int32_t), st));
                SelectParams sp{};
                sp.in_scores = tp.cand_scores;
                sp.in_ids = tp.cand_ids;
                sp.in_counts = tp.cand_counts;
                sp.n_lists = n_lists;
                sp.L = 32;
                sp.n_sel = n_sel;
                sp.list_stride_rows = rows_pad;
                sp.rows = rows_dev;
                sp.k_out = k_out;
                sp.k0 = k0;
                sp.kp = kp;
                sp.out_ids = o_ids;
                sp.out_scores = o_scores;
                sp.out_counts = o_counts;
                sp.subjects = sub32;
                sp.row_map = rowmap;
                sp.objects = E->obj32_ptr;
                sp.obj_norms = norms;
                sp.d = d;
                sp.k_cand = kc;
                sp.row_exp = E->row_exp.as<int32_t>();
                sp.obj_exp = E->obj_exp;
                const double rho = bf16 ? 0.001953125 /*2^-9*/ : 0.00048828125 /*2^-11*/;
                sp.eps_rel = (float)(2.0 * rho + rho * rho + (double)E->d_pad * 4.76837158e-7 /*2^-21*/ +
                                     std::sqrt((double)d) * 1.4551915e-11 /*2^-36*/);
                sp.max_obj_norm = E->max_obj_norm;
                sp.fb_count = fb_count;
                sp.fb_rows = fb_list;
                const size_t sel_smem = (size_t)SEL_WARPS * d * sizeof(float);
                select_kernel<true><<<grid_for(n_sel, SEL_WARPS), SEL_WARPS * 32, sel_smem, st>>>(sp);
                CK(cudaGetLastError());
                S.n_launches++;
                CK(cudaMemcpyAsync(E->h_pinned, fb_count, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
                CK(cudaStreamSynchronize(st));
                if (timed) {
                    float ms = 0.f;
                    CK(cudaEventElapsedTime(&ms, E->ev[2], E->ev[3]));
                    S.ms_main += ms;
                }
                return (int64_t)E->h_pinned[0];
            };

            if (!use_tc) {
                S.path = 0;
                run_exact(nullptr, n_rows, true, 0, k_out);
            } else {
                S.path = 1;
                S.tc_dtype = E->tc_dtype;
                S.k_cand = k_cand;
                // two failure lists of n_rows entries + two counters
                E->fb_rows.ensure(sizeof(int32_t) * (2 * n_rows + 2));
                int32_t* fb1 = E->fb_rows.as<int32_t>();
                int32_t* fb2 = fb1 + n_rows;
                int32_t* cnt = fb2 + n_rows;
                // k <= 24: one pass.  Larger k: passes of `k_pass` results; every pass is certified (or re-ranked) on its own and
                // the ids returned so far are excluded from the next pass exactly like viewed objects, so the concatenation of
                // the passes is the exact top-k in order.
                const int k_pass = k_out <= 24 ? k_out : 20;
                const int kc_pass = k_out <= 24 ? k_cand : (use_2sm ? (bf16_tc ? 30 : 25) : 32);
                if (k_out > 24) E->excl.ensure(sizeof(int32_t) * (size_t)n_rows * k_out);
                int64_t total_fb = 0, total_exact = 0;
                for (int k0 = 0; k0 < k_out; k0 += k_pass) {
                    const int kp = std::min(k_pass, k_out - k0);
                    const int kc = std::min(32, std::max(kc_pass - (k_pass - kp), kp));
                    if (k0 > 0) {
                        build_exclusion_kernel<<<grid_for(n_rows * 32, 256), 256, 0, st>>>(o_ids, n_rows, k_out, k0, (int32_t)E->id_offset,
                                                                                          E->excl.as<int32_t>());
                        CK(cudaGetLastError());
                        S.n_launches++;
                    }
                    int32_t* f1 = fb1;
                    int64_t n_fb = run_tc(nullptr, n_rows, kc, k0, kp, f1, cnt, k0 == 0);
                    total_fb += n_fb;
                    // second chance for rows whose certificate failed: same pass with the widest candidate lists (32), which
                    // only near-exact ties survive; whatever is left goes to the exhaustive fp64 kernel
                    if (n_fb > 0 && kc < 32) {
                        n_fb = run_tc(f1, n_fb, 32, k0, kp, fb2, cnt + 1, false);
                        f1 = fb2;
                    }
                    total_exact += n_fb;
                    if (n_fb > 0) {
                        const int tc_splits = S.n_splits;
                        run_exact(f1, n_fb, false, k0, k0 + kp);
                        S.n_splits = tc_splits;  // report the splits of the main kernel, not of the re-rank
                    }
                }
                S.n_fallback_rows += total_fb;
                S.n_exact_rows += total_exact;
            }

            if (E->id_offset != 0) {
                add_offset_kernel<<<grid_for(n_rows * k_out, 256), 256, 0, st>>>(o_ids, n_rows * k_out, (int32_t)E->id_offset);
                CK(cudaGetLastError());
                S.n_launches++;
            }
            if (!use_tc) {  // (the tensor-core path reads its kernel time after the synchronisation of every pass)
                CK(cudaStreamSynchronize(st));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, E->ev[2], E->ev[3]));
                S.ms_main += ms;
            }
        };

        // ---------------- chunk pipeline
        int64_t chunk = n_rows;
        if (!in_dev && use_tc) {
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
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int64_t r0 = c * chunk, r1 = std::min(n_rows, r0 + chunk);
            if (c + 1 < n_chunks) {
                stage_rows(r1, std::min(n_rows, r1 + chunk), cs);
                CK(cudaEventRecord(E->evp[(c + 1) & 1], cs));
            }
            if (n_chunks > 1) CK(cudaStreamWaitEvent(st, E->evp[c & 1], 0));
            compute_rows(r1 - r0, (sub32 && !rowmap) ? sub32 + r0 * d : sub32, rowmap ? rowmap + r0 : nullptr, indptr ? indptr + r0 : nullptr,
                         o_ids + r0 * k_out, o_scores + r0 * k_out, o_counts + r0);
            if (c + 1 == n_chunks) CK(cudaEventRecord(E->ev[4], st));
            if (!out_dev) {
                if (n_chunks > 1) {
                    CK(cudaEventRecord(E->evp[2], st));
                    CK(cudaStreamWaitEvent(cs, E->evp[2], 0));
                }
                CK(cudaMemcpyAsync(q->out_ids + r0 * k_out, o_ids + r0 * k_out, sizeof(int32_t) * (r1 - r0) * k_out, cudaMemcpyDeviceToHost, cs));
                CK(cudaMemcpyAsync(q->out_scores + r0 * k_out, o_scores + r0 * k_out, sizeof(float) * (r1 - r0) * k_out, cudaMemcpyDeviceToHost, cs));
                CK(cudaMemcpyAsync(q->out_counts + r0, o_counts + r0, sizeof(int32_t) * (r1 - r0), cudaMemcpyDeviceToHost, cs));
                S.d2h_bytes += (int64_t)((r1 - r0) * k_out * 8 + (r1 - r0) * 4);
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
        CK(cudaEventElapsedTime(&S.ms_d2h, E->ev[4], E->ev[5]));  // exposed part: the last chunk's results
    } catch (const CudaError& ce) {
        return fail(ce.e == cudaErrorMemoryAllocation ? B200_E_NOMEM : B200_E_CUDA, "b200_rank_topk: %s failed at line %d: %s", ce.what,
                    ce.line, cudaGetErrorString(ce.e));
    }
    if (stats) *stats = S;
    return B200_OK;
}

int b200_rank_merge(int32_t device, void* stream, int32_t n_lists, int64_t n_rows, int32_t k, const int32_t* ids,
                    const float* scores, const int32_t* counts, int32_t* out_ids, float* out_scores, int32_t* out_counts) {
    if (n_lists <= 0 || n_rows < 0 || k <= 0 || !ids || !scores || !counts || !out_ids || !out_scores || !out_counts)
        return fail(B200_E_INVALID, "b200_rank_merge: bad arguments");
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
            sp.n_lists = n_lists;
            sp.L = k;
            sp.n_sel = n_rows;
            sp.list_stride_rows = n_rows;
            sp.k_out = k;
            sp.k0 = k0;
            sp.kp = std::min(32, k - k0);
            sp.out_ids = out_ids;
            sp.out_scores = out_scores;
            sp.out_counts = out_counts;
            select_kernel<false><<<grid_for(n_rows, SEL_WARPS), SEL_WARPS * 32, 0, st>>>(sp);
            CK(cudaGetLastError());
        }
    } catch (const CudaError& ce) {
        return fail(B200_E_CUDA, "b200_rank_merge: %s failed: %s", ce.what, cudaGetErrorString(ce.e));
    }
    return B200_OK;
}

}  // extern "C"
