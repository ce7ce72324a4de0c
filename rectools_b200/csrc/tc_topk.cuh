// Fused tensor-core scoring + streaming top-K' candidate selection for sm_100a (tcgen05 / TMEM / TMA).
//
// One persistent CTA per SM.  A work item = (subject tile of S*128 rows, object split).  The subject tile stays
// resident in shared memory; object tiles of 128 rows are streamed through a TMA -> mbarrier ring in 64-column
// (128-byte, SWIZZLE_128B) blocks; `tcgen05.mma` (M=128, N=128, K=16, fp16/bf16 -> fp32) accumulates each
// [128 x 128] score tile in TMEM (2 buffers x S sub-tiles x 128 columns = up to all 512 columns); the epilogue
// warps read the accumulators back with `tcgen05.ld` (thread = subject row, 32 consecutive objects per load),
// compare against the row's running K'-th best score and only on a hit (rare after warm-up) look the object up in
// the row's `filter_pairs_csr` slice and insert it into the row's sorted candidate list (shared memory, one
// list entry per lane).  Score rows never reach HBM: only K' (score, id) pairs per row and split are written.
//
// This replaces `scores = query @ items.T` + mask + select of implicit's top-k (call site
// rectools/models/rank/rank_implicit.py:264-272, :175-182) and `TorchRanker.rank`'s batched matmul / masked_fill /
// torch.topk (rectools/models/rank/rank_torch.py:133-152) as a CANDIDATE generator; exact scores and the final
// order come from select_kernel<true> (select.cuh).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace b200 {
namespace tc {

constexpr int TILE_M = 128;              // subject rows per sub-tile = TMEM lanes
constexpr int TILE_N = 128;              // objects per tile = TMEM columns per accumulator
constexpr int KBLK = 64;                 // 16-bit elements per shared-memory block row (128 B, one swizzle atom)
constexpr int BLK_BYTES = 128 * KBLK * 2;  // 16 KiB: [128 rows][128 B]
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 384;         // warp 0: TMA, warp 1: MMA, warp 2: TMEM alloc, warp 3: idle, warps 4-11: epilogue
constexpr int EPI_WARP0 = 4;
constexpr int MAX_STAGES = 12;
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 232448;       // 227 KiB

struct TcParams {
    int32_t s_sub;            // sub-tiles per CTA (1 or 2)
    int32_t kblocks;          // d_pad / 64
    int32_t n_stages;         // object ring depth (blocks of 16 KiB)
    int32_t k_cand;           // K' <= 32
    int64_t n_rows;           // valid subject rows
    int64_t n_pos;            // valid object positions
    int32_t n_row_tiles;
    int32_t n_splits;
    int32_t n_obj_tiles;
    int32_t tiles_per_split;
    uint32_t idesc;           // UMMA instruction descriptor
    const int32_t* pos2obj;   // nullable whitelist map
    const int64_t* indptr;    // nullable CSR filter by subject row
    const int32_t* indices;
    const int32_t* row_ids;   // nullable: batch row -> row of the CSR filter (re-ranked subsets)
    const int32_t* excl;      // nullable: [rows][excl_stride] ids already returned by earlier passes (k > 24), sorted ascending
    int32_t excl_stride;
    int32_t excl_n;           // ids per row in `excl` (rows with fewer results are padded with B200_PAD_ID)
    int32_t id_off;           // global id = local object id + id_off (CSR column ids are global)
    float* cand_scores;       // [n_splits][rows_pad][32]
    int32_t* cand_ids;
    int32_t* cand_counts;     // [n_splits][rows_pad]
    int64_t rows_pad;
    int32_t debug_mode;       // 0 = normal; 1 = no candidates (fast path only); 2 = epilogue skips the TMEM reads (measurement hooks)
    // carousel (2-SM kernel): a work item starts streaming the objects where the other CTA pairs currently are, so that
    // all pairs keep reading the same few MB of the object matrix and the L2 serves 73 of 74 reads (nullptr: start at t0)
    int32_t* front;           // [n_splits] object tile most recently issued by some pair
    int32_t* starts;          // [n_pairs][starts_stride] start tile chosen for each work item (-1: not decided yet)
    int32_t starts_stride;
};

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Wait for the phase with the given parity to complete.  A watchdog turns a protocol bug into a trap, not a hang
// (a failed try_wait already suspends the thread for a few hundred cycles, so 2^24 failures are seconds).
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > (1u << 24)) __trap();
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tiled TMA load global -> shared, completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// One lane of the (converged) warp; the same lane every time, so tcgen05.commit tracks the MMAs it issued.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, 128 x N x 16, issued by one thread.  The two shared-memory matrix descriptors
// differ only in their low word (start address >> 4); the high word (stride, version, swizzle) is shared.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                         uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accum)
        : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i), columns [c, c+32).
// The load is asynchronous: the registers are valid only after tmem_ld_wait() on the same array.
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// Wait for every outstanding tcgen05.ld of this thread.  The registers are threaded through the asm as in/out
// operands so that the compiler cannot schedule their first use above the wait.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}

// Synchronous wide TMEM reads (load + wait in one asm statement so that no use can be scheduled in between):
// thread i of the warp gets TMEM lane (base_lane + i), 64 / 128 consecutive fp32 columns.
#define B200_R8(a, n) "=r"(a[n]), "=r"(a[n + 1]), "=r"(a[n + 2]), "=r"(a[n + 3]), "=r"(a[n + 4]), "=r"(a[n + 5]), "=r"(a[n + 6]), "=r"(a[n + 7])
__device__ __forceinline__ void tmem_ld64_sync(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : B200_R8(r, 0), B200_R8(r, 8), B200_R8(r, 16), B200_R8(r, 24), B200_R8(r, 32), B200_R8(r, 40), B200_R8(r, 48),
          B200_R8(r, 56)
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld128_sync(uint32_t taddr, uint32_t (&r)[128]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x128.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63, "
        "%64, %65, %66, %67, %68, %69, %70, %71, %72, %73, %74, %75, %76, %77, %78, %79, "
        "%80, %81, %82, %83, %84, %85, %86, %87, %88, %89, %90, %91, %92, %93, %94, %95, "
        "%96, %97, %98, %99, %100, %101, %102, %103, %104, %105, %106, %107, %108, %109, %110, %111, "
        "%112, %113, %114, %115, %116, %117, %118, %119, %120, %121, %122, %123, %124, %125, %126, %127}, [%128];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : B200_R8(r, 0), B200_R8(r, 8), B200_R8(r, 16), B200_R8(r, 24), B200_R8(r, 32), B200_R8(r, 40), B200_R8(r, 48),
          B200_R8(r, 56), B200_R8(r, 64), B200_R8(r, 72), B200_R8(r, 80), B200_R8(r, 88), B200_R8(r, 96), B200_R8(r, 104),
          B200_R8(r, 112), B200_R8(r, 120)
        : "r"(taddr)
        : "memory");
}

// Shared-memory matrix descriptor of a K-major operand block: 128-byte rows, SWIZZLE_128B, 8-row groups 1024 B apart.
//   lo: bits [0,14) start address >> 4, bits [16,30) leading byte offset >> 4 (unused for swizzled K-major: 0)
//   hi: bits [0,14) stride byte offset >> 4 (1024 >> 4), bits [14,16) descriptor version 1 (sm_100), bits [29,32) layout 2
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t saddr) { return (saddr & 0x3FFFF) >> 4; }
constexpr uint32_t SMEM_DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);

__device__ __forceinline__ float max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// ------------------------------------------------------------------------------------------------ epilogue helpers
// Per-row state of the streaming selection (registers of the thread that owns the row).
struct RowState {
    float thr;    // smallest score of the row's candidate list once it holds K' entries (-inf before, +inf: padded row)
    int cnt;      // entries in the list (<= K')
    int minpos;   // slot of the smallest entry once the list is full
    int nv;       // next viewed global object id >= the stream position (B200_PAD_ID when the CSR row is exhausted)
    int64_t cur;  // index of `nv` in csr indices
    int64_t fhi;  // end of the row's CSR slice
    // multi-pass ranking (k > 24): objects returned by earlier passes are excluded the same way
    const int32_t* xrow;  // this row's sorted exclusion list (nullptr: none)
    int xcur, xnv;        // cursor / next excluded global object id
};

// Candidate lists live in shared memory as [slot][lane]: the thread that owns a row reads and writes only its own
// column (bank = lane, conflict-free), so all 32 rows of a warp can take candidates at the same time.
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ int lds_s32(uint32_t a) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_s32(uint32_t a, int v) { asm volatile("st.shared.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Keep the K' best (score, id) pairs of a row: append while the list is short, afterwards overwrite the current
// minimum and re-scan for the new one (K' independent shared-memory loads; runs for all rows of the warp in parallel).
// ls / li = shared addresses of slot 0 of this thread's column in the score / id arrays.
__device__ __forceinline__ void list_insert(uint32_t ls, uint32_t li, int kc, RowState& rs, float val, int obj) {
    const int slot = rs.cnt < kc ? rs.cnt : rs.minpos;
    sts_f32(ls + slot * 128, val);
    sts_s32(li + slot * 128, obj);
    if (rs.cnt < kc && ++rs.cnt < kc) return;
    float mn = INFINITY;
    int mp = 0;
#pragma unroll 8
    for (int e = 0; e < kc; ++e) {
        const float x = lds_f32(ls + e * 128);
        if (x < mn) {
            mn = x;
            mp = e;
        }
    }
    rs.minpos = mp;
    rs.thr = fmaxf(rs.thr, mn);  // never loosen a bound borrowed from the row's other list
}

// Objects are visited in ascending id order, so the filter_pairs_csr lookup is a merge, not a search: `nv` trails the
// stream and is advanced only when a candidate passes it (csr_advance_coop below).

// Same merge cursor over the (short, <= k entries) list of objects already returned by earlier passes of a k > 24 query.
__device__ __forceinline__ bool is_excluded(RowState& rs, int n, int g) {
    while (rs.xnv < g) {
        ++rs.xcur;
        rs.xnv = rs.xcur < n ? __ldg(rs.xrow + rs.xcur) : B200_PAD_ID;
    }
    return rs.xnv == g;
}

// Position the two cursors of a row at the first object (global id g_first) of a work item.
__device__ __forceinline__ void row_cursors_init(const TcParams& p, RowState& rs, int64_t frow, int g_first) {
    rs.nv = B200_PAD_ID;
    rs.cur = 0;
    rs.fhi = 0;
    rs.xrow = nullptr;
    rs.xcur = 0;
    rs.xnv = B200_PAD_ID;
    if (frow < 0) return;
    if (p.indptr) {
        int64_t lo = p.indptr[frow];
        rs.fhi = p.indptr[frow + 1];
        int64_t hi = rs.fhi;
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (__ldg(p.indices + mid) < g_first)
                lo = mid + 1;
            else
                hi = mid;
        }
        rs.cur = lo;
        rs.nv = lo < rs.fhi ? __ldg(p.indices + lo) : B200_PAD_ID;
    }
    if (p.excl) {
        rs.xrow = p.excl + frow * p.excl_stride;
        int lo = 0, hi = p.excl_n;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(rs.xrow + mid) < g_first)
                lo = mid + 1;
            else
                hi = mid;
        }
        rs.xcur = lo;
        rs.xnv = lo < p.excl_n ? __ldg(rs.xrow + lo) : B200_PAD_ID;
    }
}

// v[j] for a run-time j without local memory: 5-level select tree (31 SEL), cheaper than spilling the chunk.
__device__ __forceinline__ float select32(const float (&v)[32], int j) {
    float a[16], b[8], c[4], d[2];
    const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8, b4 = j & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b0 ? v[2 * i + 1] : v[2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = b1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = b2 ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = b3 ? c[2 * i + 1] : c[2 * i];
    return b4 ? d[1] : d[0];
}

// Slow path of one chunk (compact on purpose: the unrolled per-column version thrashed the instruction cache).
// Every lane builds the bit mask of its columns above the row threshold and then drains it: lowest pending column
// first, re-checked against the possibly raised threshold, padded / viewed objects dropped, survivors inserted
// into the row's candidate list by the thread that owns the row.
template <int J0, int J1>
__device__ __forceinline__ unsigned hit_mask(const float (&v)[32], float thr) {
    unsigned m = 0;
#pragma unroll
    for (int j = J0; j < J1; ++j) m |= (v[j] > thr) ? (1u << j) : 0u;
    return m;
}

// Warp-cooperative advance of ONE row's CSR cursor (the row owned by lane `src`) to the first viewed id >= that lane's
// candidate id: all 32 lanes read 32 consecutive column ids of the row's slice at once, so a hit costs one global
// round trip instead of a chain of dependent loads (measured bottleneck of the per-lane search: in the sparse tail of the
// stream a chunk's slow path is entered for a single lane, whose 2 linear steps + lower_bound were ~9 dependent L2
// reads = several tile times).  Long slices are first narrowed by a 32-ary search.
__device__ __forceinline__ void csr_advance_coop(const int32_t* __restrict__ indices, RowState& rs, int g, int src, int lane) {
    const int gs = __shfl_sync(B200_FULL_MASK, g, src);
    long long base = __shfl_sync(B200_FULL_MASK, (long long)rs.cur, src) + 1;  // entries up to `cur` are < gs
    const long long fhi = __shfl_sync(B200_FULL_MASK, (long long)rs.fhi, src);
    int x, adv;
    bool narrowed = false;
    for (;;) {
        const long long i = base + lane;
        x = i < fhi ? __ldg(indices + i) : B200_PAD_ID;
        adv = __popc(__ballot_sync(B200_FULL_MASK, x < gs));  // sorted slice: the lanes below gs form a prefix
        base += adv;
        if (adv < 32) break;
        if (!narrowed && fhi - base > 64) {
            // 32-ary narrowing: afterwards every entry before `base` is < gs and the answer lies within 32 entries
            long long hi = fhi;
            while (hi - base > 32) {
                const long long step = (hi - base + 31) >> 5;
                const long long pi = base + (long long)lane * step;
                const bool below = pi < hi && __ldg(indices + pi) < gs;
                const int c = __popc(__ballot_sync(B200_FULL_MASK, below));
                if (c == 0) break;  // indices[base] >= gs
                const long long nb = base + (long long)(c - 1) * step + 1;
                if (c < 32 && base + (long long)c * step < hi) hi = base + (long long)c * step;
                base = nb;
            }
            narrowed = true;
        }
    }
    const int nvn = __shfl_sync(B200_FULL_MASK, x, adv);  // lane `adv` holds the first id >= gs (PAD_ID past the end)
    if (lane == src) {
        rs.cur = base;
        rs.nv = nvn;
    }
}

__device__ __forceinline__ void scan_chunk(const float (&v)[32], float g0, float g1, float g2, float g3, int64_t pos0,
                                           const TcParams& p, uint32_t ls, uint32_t li, int kc, RowState& rs) {
    const int lane = threadIdx.x & 31;
    // per-lane bit mask of the columns above the row threshold, built only for the column groups whose maximum
    // (already known from the fast path) shows a hit somewhere in the warp
    unsigned hits = 0;
    if (__any_sync(B200_FULL_MASK, g0 > rs.thr)) hits |= hit_mask<0, 9>(v, rs.thr);
    if (__any_sync(B200_FULL_MASK, g1 > rs.thr)) hits |= hit_mask<9, 18>(v, rs.thr);
    if (__any_sync(B200_FULL_MASK, g2 > rs.thr)) hits |= hit_mask<18, 27>(v, rs.thr);
    if (__any_sync(B200_FULL_MASK, g3 > rs.thr)) hits |= hit_mask<27, 32>(v, rs.thr);
    // drain: every thread works through its own columns in ascending object order (what the CSR cursor needs);
    // rows are independent, so all lanes insert concurrently; only the (rare) cursor advances are done lane by lane
    // with the whole warp helping
    while (__any_sync(B200_FULL_MASK, hits != 0)) {
        bool cand = false;
        float val = 0.f;
        int obj = 0, g = 0;
        if (hits) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            val = select32(v, j);
            const int64_t pos = pos0 + j;
            if (val > rs.thr && pos < p.n_pos) {  // the threshold may have risen since the mask was built
                obj = p.pos2obj ? __ldg(p.pos2obj + pos) : (int)pos;
                g = obj + p.id_off;
                cand = true;
            }
        }
        unsigned need = __ballot_sync(B200_FULL_MASK, cand && rs.nv < g);
        while (need) {
            const int src = __ffs(need) - 1;
            need &= need - 1;
            csr_advance_coop(p.indices, rs, g, src, lane);
        }
        if (cand && rs.nv != g && !(rs.xrow && is_excluded(rs, p.excl_n, g))) list_insert(ls, li, kc, rs, val, obj);
    }
}

// One 32-column chunk of one accumulator row per thread: 3-input max tree against the row threshold (fast path,
// ~0.6 instructions per score); the scan above runs only when some row of the warp has a hit.
template <int OFF = 0, int NREG = 32>
__device__ __forceinline__ void process_chunk(const uint32_t (&r)[NREG], int64_t pos0, const TcParams& p, uint32_t ls,
                                              uint32_t li, int kc, RowState& rs) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[OFF + i]);
    const float g0 = max3(max3(v[0], v[1], v[2]), max3(v[3], v[4], v[5]), max3(v[6], v[7], v[8]));
    const float g1 = max3(max3(v[9], v[10], v[11]), max3(v[12], v[13], v[14]), max3(v[15], v[16], v[17]));
    const float g2 = max3(max3(v[18], v[19], v[20]), max3(v[21], v[22], v[23]), max3(v[24], v[25], v[26]));
    const float g3 = max3(max3(v[27], v[28], v[29]), v[30], v[31]);
    const float mx = fmaxf(max3(g0, g1, g2), g3);
    if (__any_sync(B200_FULL_MASK, mx > rs.thr)) scan_chunk(v, g0, g1, g2, g3, pos0, p, ls, li, kc, rs);
}

// ------------------------------------------------------------------------------------------------ the kernel
__global__ void __launch_bounds__(NUM_THREADS, 1)
tc_topk_kernel(const __grid_constant__ CUtensorMap tm_sub, const __grid_constant__ CUtensorMap tm_obj, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int S = p.s_sub, KB = p.kblocks, NS = p.n_stages;
    uint8_t* sA = smem;                                      // [S][KB] blocks
    uint8_t* sB = sA + (size_t)S * KB * BLK_BYTES;           // [NS] blocks
    float* sLs = reinterpret_cast<float*>(sB + (size_t)NS * BLK_BYTES);  // [S*128][32] candidate scores
    int* sLi = reinterpret_cast<int*>(sLs + S * TILE_M * 32);            // [S*128][32] candidate ids
    uint64_t* bars = reinterpret_cast<uint64_t*>(sLi + S * TILE_M * 32);
    // barrier map: full[0..MAX), empty[MAX..2MAX), a_full, a_empty, t_full[2], t_empty[2]
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + MAX_STAGES);
    const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES);
    const uint32_t bar_aempty = smem_u32(bars + 2 * MAX_STAGES + 1);
    const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES + 2);
    const uint32_t bar_tempty = smem_u32(bars + 2 * MAX_STAGES + 4);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(bar_full + 8 * i, 1);
            mbar_init(bar_empty + 8 * i, 1);
        }
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 4 * S);  // one arrive per epilogue warp
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_sub);
        tma_prefetch_desc(&tm_obj);
    }
    if (warp == 2) {
        tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();  // the CTA allocated all 512 columns, so its TMEM window starts at lane 0 / column 0

    const int n_work = p.n_row_tiles * p.n_splits;

    if (warp == 0) {
        // ===================================================================== TMA producer (converged warp, elected lane issues)
        uint32_t stage = 0, ph = 0, work_it = 0;
        const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++work_it) {
            const int split = w / p.n_row_tiles, rt = w - split * p.n_row_tiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
            if (work_it > 0) mbar_wait(bar_aempty, (work_it - 1) & 1);  // previous tile's MMAs are done with sA
            if (elect_one()) {
                mbar_arrive_expect_tx(bar_afull, (uint32_t)(S * KB * BLK_BYTES));
                for (int s = 0; s < S; ++s)
                    for (int kb = 0; kb < KB; ++kb)
                        tma_load_2d(sA_u + (uint32_t)(s * KB + kb) * BLK_BYTES, &tm_sub, bar_afull, kb * KBLK,
                                    (rt * S + s) * TILE_M);
            }
            __syncwarp();
            for (int t = t0; t < t1; ++t) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(bar_full + 8 * stage, BLK_BYTES);
                        tma_load_2d(sB_u + stage * BLK_BYTES, &tm_obj, bar_full + 8 * stage, kb * KBLK, t * TILE_N);
                    }
                    __syncwarp();
                    if (++stage == (uint32_t)NS) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer
        // The whole warp walks the loops converged and one elected lane issues: every operand of tcgen05.mma is then
        // provably warp-uniform (loop counters, kernel parameters, the dynamic-smem base; TMEM base 0 because the CTA
        // owns all 512 columns).  Issuing from inside an `if (lane == 0)` made the compiler wrap every UTCHMMA in an
        // elect/broadcast retry loop that cost ~170 cycles per instruction (measured: tensor pipe 36 % busy).
        uint32_t stage = 0, ph = 0, tile_it = 0, work_it = 0;
        const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
        constexpr uint32_t BLK16 = BLK_BYTES >> 4;  // one 16 KiB block in descriptor address units
        const uint32_t a_sub1 = (uint32_t)KB * BLK16;  // second sub-tile of the subject tile
        const uint32_t idesc = p.idesc;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x, ++work_it) {
            const int split = w / p.n_row_tiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
            mbar_wait(bar_afull, work_it & 1);
            tc_fence_after();
            for (int t = t0; t < t1; ++t, ++tile_it) {
                const uint32_t buf = tile_it & 1, tph = (tile_it >> 1) & 1;
                mbar_wait(bar_tempty + 8 * buf, tph ^ 1);  // epilogue has drained this accumulator pair
                tc_fence_after();
                const uint32_t d0 = buf * (uint32_t)(S * TILE_N);
                uint32_t a_lo = a_lo0;
                for (int kb = 0; kb < KB; ++kb, a_lo += BLK16) {
                    mbar_wait(bar_full + 8 * stage, ph);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t b_lo = b_lo0 + stage * BLK16;
                        // +32 B per K step inside the 128 B swizzle atom = +2 in descriptor address units
                        umma_f16(d0, a_lo, b_lo, SMEM_DESC_HI, idesc, (uint32_t)(kb != 0));
                        umma_f16(d0, a_lo + 2, b_lo + 2, SMEM_DESC_HI, idesc, 1u);
                        umma_f16(d0, a_lo + 4, b_lo + 4, SMEM_DESC_HI, idesc, 1u);
                        umma_f16(d0, a_lo + 6, b_lo + 6, SMEM_DESC_HI, idesc, 1u);
                        if (S == 2) {
                            const uint32_t d1 = d0 + TILE_N, a1 = a_lo + a_sub1;
                            umma_f16(d1, a1, b_lo, SMEM_DESC_HI, idesc, (uint32_t)(kb != 0));
                            umma_f16(d1, a1 + 2, b_lo + 2, SMEM_DESC_HI, idesc, 1u);
                            umma_f16(d1, a1 + 4, b_lo + 4, SMEM_DESC_HI, idesc, 1u);
                            umma_f16(d1, a1 + 6, b_lo + 6, SMEM_DESC_HI, idesc, 1u);
                        }
                        umma_commit(bar_empty + 8 * stage);  // ring slot is free once these MMAs retire
                        if (kb == KB - 1) umma_commit(bar_tfull + 8 * buf);  // accumulators of this tile are complete
                    }
                    if (++stage == (uint32_t)NS) {
                        stage = 0;
                        ph ^= 1;
                    }
                }
            }
            if (elect_one()) umma_commit(bar_aempty);
        }
    } else if (warp >= EPI_WARP0 && warp < EPI_WARP0 + 4 * S) {
        // ===================================================================== epilogue: select candidates
        const int ew = warp - EPI_WARP0;
        const int s = ew >> 2, quarter = ew & 3;  // quarter == warp % 4: the TMEM lanes this warp may read
        const int wrow0 = s * TILE_M + quarter * 32;  // first CTA-local row of this warp
        const uint32_t ls = smem_u32(sLs + (size_t)wrow0 * 32) + lane * 4;  // [slot][lane] arrays of this warp
        const uint32_t li = smem_u32(sLi + (size_t)wrow0 * 32) + lane * 4;
        const int kc = p.k_cand;
        uint32_t tile_it = 0;
        for (int w = blockIdx.x; w < n_work; w += gridDim.x) {
            const int split = w / p.n_row_tiles, rt = w - split * p.n_row_tiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
            const int64_t grow0 = (int64_t)rt * S * TILE_M + wrow0;  // global row of lane 0
            const int64_t grow = grow0 + lane;
            const bool row_ok = grow < p.n_rows;
            RowState rs;
            rs.thr = (row_ok && p.debug_mode == 0) ? -INFINITY : INFINITY;  // padded rows never produce candidates
            rs.cnt = 0;
            rs.minpos = 0;
            {
                const int64_t pos_first = (int64_t)t0 * TILE_N;
                const bool live = row_ok && pos_first < p.n_pos;
                const int g_first = live ? (p.pos2obj ? __ldg(p.pos2obj + pos_first) : (int)pos_first) + p.id_off : 0;
                row_cursors_init(p, rs, live ? (p.row_ids ? (int64_t)p.row_ids[grow] : grow) : -1, g_first);
            }
            for (int t = t0; t < t1; ++t, ++tile_it) {
                const uint32_t buf = tile_it & 1, tph = (tile_it >> 1) & 1;
                mbar_wait(bar_tfull + 8 * buf, tph);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)((buf * S + s) * TILE_N);
                const int64_t pos_t = (int64_t)t * TILE_N;
                // software-pipelined TMEM reads: chunk c+1 is in flight while chunk c is scanned
                if (p.debug_mode == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                    continue;
                }
                uint32_t ra[32], rb[32];
                tmem_ld_issue(tbase, ra);
#pragma unroll 1
                for (int h = 0; h < 2; ++h) {  // two chunk pairs; not unrolled to keep the code I-cache resident
                    tmem_ld_wait(ra);
                    tmem_ld_issue(tbase + h * 64 + 32, rb);
                    process_chunk(ra, pos_t + h * 64, p, ls, li, kc, rs);
                    tmem_ld_wait(rb);
                    if (h == 0) {
                        tmem_ld_issue(tbase + 64, ra);
                    } else {
                        // every TMEM read of this tile has completed: hand the accumulator back before the last scan
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_tempty + 8 * buf);
                    }
                    process_chunk(rb, pos_t + h * 64 + 32, p, ls, li, kc, rs);
                }
            }
            // ---- write this thread's candidate list (unsorted; select_kernel<true> orders it)
            if (row_ok) {
                const int64_t lrow = (int64_t)split * p.rows_pad + grow;
                for (int e = 0; e < 32; ++e) {
                    const bool keep = e < rs.cnt;
                    p.cand_scores[lrow * 32 + e] = keep ? lds_f32(ls + e * 128) : -INFINITY;
                    p.cand_ids[lrow * 32 + e] = keep ? lds_s32(li + e * 128) : B200_PAD_ID;
                }
                p.cand_counts[lrow] = rs.cnt;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace b200
