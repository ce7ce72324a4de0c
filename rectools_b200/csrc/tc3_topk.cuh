// Third generation of the fused scoring + candidate-selection kernel (default): the 2-SM MMA / TMA pipeline of
// tc2_topk.cuh (CTA pairs, tcgen05.mma.cta_group::2, 256 x 256 x 16, two 256-column accumulators per CTA) with a new epilogue.
//
// What the source-level ncu profile of the previous epilogue showed (profiles/r01_ncu_source_*.txt): the epilogue warps
// are latency bound (2 warps per scheduler, ~0.15 instructions per cycle and warp), ~330 warp instructions per tile
// at N = 1M of which ~150 were per-tile bookkeeping (spill reloads of loop invariants, re-materialised addresses,
// generic-address volatile accesses), ~80 the threshold scan and ~100 the candidate slow path -- a slow path that ran
// once per hit with 31 of 32 lanes idle in the sparse tail of the object stream.  This kernel
//   * gives the epilogue warps 232 registers (setmaxnreg: the TMA / MMA warp group keeps 40), so the staged accumulator
//     slice (128 registers) and the row state stay in registers and nothing is spilled;
//   * scans the whole 128-column slice against the row threshold with ONE vote per tile;
//   * on a hit only EXTRACTS it (column mask, value select) into a small per-thread FIFO in shared memory; the
//     expensive part -- re-check against the current threshold, filter_pairs_csr lookup, insertion into the row's
//     candidate list -- runs as single bounded steps (the oldest pending hit of all 32 rows of the warp at once, at most
//     one step per tile, by default every 16th tile or as soon as some row has four hits waiting: batching the steps is
//     worth 7 %), and the CSR lookup never waits for memory: a row's viewed ids are read through a four-entry window
//     whose loads are issued one step before they are needed (measured: draining whole FIFOs at once, or advancing the
//     CSR cursors one row after the other, stalls the accumulator hand-over for tens of tile times);
//   * keeps per-tile bookkeeping incremental (tile index, accumulator parity, object position).
// Deferred hits only ever see a threshold that is older (lower) than the current one, i.e. the filter is weaker, never
// wrong; every FIFO is emptied before the row's CSR window is repositioned and before its list is written out.
//
// Replaces the same reference code as tc_topk.cuh (rank_implicit.py:264-272 / rank_torch.py:133-152) as a candidate
// generator; final scores / order / certificate come from select_kernel<true>.
#pragma once
#include "tc2_topk.cuh"

namespace b200 {
namespace tc {

// Warp group 0 (warps 0..3): TMA producer, MMA issuer + TMEM allocator, two idle warps; warp groups 1 and 2 (warps 4..11):
// epilogue.  An SM sub-partition (16384 registers) hosts one warp of each group, so a uniform split would cap every
// thread at 168 registers; `setmaxnreg` moves the registers warp group 0 does not need to the epilogue warps.
constexpr int T3_THREADS = 384;
constexpr int T3_EPI0 = 4;
constexpr int T3_REGS_LOW = 40, T3_REGS_EPI = 232;  // 32 * (40 + 2 * 232) = 16128 <= 16384 per sub-partition
#ifndef B200_T3_STEP_LAZY
#define B200_T3_STEP_LAZY 1  // deferred work every B200_T3_STEP_PERIOD-th tile unless some row has B200_T3_BACKLOG or more hits pending
#endif
#ifndef B200_T3_STEP_PERIOD
#define B200_T3_STEP_PERIOD 16  // power of two
#endif
#ifndef B200_T3_BACKLOG
#define B200_T3_BACKLOG 4
#endif
#ifndef B200_T3_Q
#define B200_T3_Q 8  // measured together with the step period / backlog: profiles/r01_ab_tc3.txt
#endif
constexpr int T3_Q = B200_T3_Q;                    // deferred hits per thread
constexpr int T3_QSTRIDE = 8 * 32 * 8;             // bytes between FIFO slots: [slot][epilogue thread] x (score, position)
constexpr int T3_QBYTES = T3_Q * T3_QSTRIDE;       // 16 KiB per CTA at 8 slots
constexpr int T3_TN = 256, T3_HALF = 128;

__device__ __forceinline__ void sts_v2(uint32_t a, float x, uint32_t y) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(__float_as_uint(x)), "r"(y) : "memory");
}
__device__ __forceinline__ void lds_v2(uint32_t a, float& x, uint32_t& y) {
    uint32_t xb;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(xb), "=r"(y) : "r"(a) : "memory");
    x = __uint_as_float(xb);
}
// (work-item tag, threshold) pairs exchanged between the two threads that scan the two column halves of a row
__device__ __forceinline__ void sts_thr(uint32_t a, uint32_t tag, float thr) {
    asm volatile("{\n\t.reg .b64 t;\n\tmov.b64 t, {%1, %2};\n\tst.volatile.shared.b64 [%0], t;\n\t}"
                 ::"r"(a), "r"(__float_as_uint(thr)), "r"(tag)
                 : "memory");
}
__device__ __forceinline__ void lds_thr(uint32_t a, uint32_t& tag, float& thr) {
    uint32_t tb;
    asm volatile("{\n\t.reg .b64 t;\n\tld.volatile.shared.b64 t, [%2];\n\tmov.b64 {%0, %1}, t;\n\t}"
                 : "=r"(tb), "=r"(tag)
                 : "r"(a)
                 : "memory");
    thr = __uint_as_float(tb);
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

template <int N>
__device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ float fu(uint32_t x) { return __uint_as_float(x); }

// Pin a loop-invariant value in a register: without this the compiler re-derives shared-memory / TMEM addresses from
// %tid, the shared window base and the kernel parameters in every tile iteration (~45 instructions per tile in the
// previous kernel) instead of spending a register on them.
__device__ __forceinline__ uint32_t pin(uint32_t x) {
    uint32_t y;
    asm volatile("mov.u32 %0, %1;" : "=r"(y) : "r"(x));
    return y;
}

// Maximum of the 32 staged scores r[OFF .. OFF+32) (11 three-input maxima + 1).
template <int OFF, int NR>
__device__ __forceinline__ float chunk_max(const uint32_t (&r)[NR]) {
    const float g0 = max3(max3(fu(r[OFF + 0]), fu(r[OFF + 1]), fu(r[OFF + 2])), max3(fu(r[OFF + 3]), fu(r[OFF + 4]), fu(r[OFF + 5])),
                          max3(fu(r[OFF + 6]), fu(r[OFF + 7]), fu(r[OFF + 8])));
    const float g1 = max3(max3(fu(r[OFF + 9]), fu(r[OFF + 10]), fu(r[OFF + 11])), max3(fu(r[OFF + 12]), fu(r[OFF + 13]), fu(r[OFF + 14])),
                          max3(fu(r[OFF + 15]), fu(r[OFF + 16]), fu(r[OFF + 17])));
    const float g2 = max3(max3(fu(r[OFF + 18]), fu(r[OFF + 19]), fu(r[OFF + 20])), max3(fu(r[OFF + 21]), fu(r[OFF + 22]), fu(r[OFF + 23])),
                          max3(fu(r[OFF + 24]), fu(r[OFF + 25]), fu(r[OFF + 26])));
    const float g3 = max3(max3(fu(r[OFF + 27]), fu(r[OFF + 28]), fu(r[OFF + 29])), fu(r[OFF + 30]), fu(r[OFF + 31]));
    return fmaxf(max3(g0, g1, g2), g3);
}

template <int OFF, int J0, int J1, int NR>
__device__ __forceinline__ unsigned group_mask(const uint32_t (&r)[NR], float thr) {
    unsigned m = 0;
#pragma unroll
    for (int j = J0; j < J1; ++j) m |= (fu(r[OFF + j]) > thr) ? (1u << j) : 0u;
    return m;
}

// Per-lane bit mask of the columns of chunk OFF above the row threshold; the mask of a 9-column group is built only
// when the group's maximum shows a hit somewhere in the warp.
template <int OFF, int NR>
__device__ __forceinline__ unsigned chunk_hits(const uint32_t (&r)[NR], float thr) {
    const float g0 = max3(max3(fu(r[OFF + 0]), fu(r[OFF + 1]), fu(r[OFF + 2])), max3(fu(r[OFF + 3]), fu(r[OFF + 4]), fu(r[OFF + 5])),
                          max3(fu(r[OFF + 6]), fu(r[OFF + 7]), fu(r[OFF + 8])));
    const float g1 = max3(max3(fu(r[OFF + 9]), fu(r[OFF + 10]), fu(r[OFF + 11])), max3(fu(r[OFF + 12]), fu(r[OFF + 13]), fu(r[OFF + 14])),
                          max3(fu(r[OFF + 15]), fu(r[OFF + 16]), fu(r[OFF + 17])));
    const float g2 = max3(max3(fu(r[OFF + 18]), fu(r[OFF + 19]), fu(r[OFF + 20])), max3(fu(r[OFF + 21]), fu(r[OFF + 22]), fu(r[OFF + 23])),
                          max3(fu(r[OFF + 24]), fu(r[OFF + 25]), fu(r[OFF + 26])));
    const float g3 = max3(max3(fu(r[OFF + 27]), fu(r[OFF + 28]), fu(r[OFF + 29])), fu(r[OFF + 30]), fu(r[OFF + 31]));
    unsigned hits = 0;
    if (__any_sync(B200_FULL_MASK, g0 > thr)) hits |= group_mask<OFF, 0, 9>(r, thr);
    if (__any_sync(B200_FULL_MASK, g1 > thr)) hits |= group_mask<OFF, 9, 18>(r, thr);
    if (__any_sync(B200_FULL_MASK, g2 > thr)) hits |= group_mask<OFF, 18, 27>(r, thr);
    if (__any_sync(B200_FULL_MASK, g3 > thr)) hits |= group_mask<OFF, 27, 32>(r, thr);
    return hits;
}

// r[OFF + j] for a run-time j without local memory: 5-level select tree (31 SEL).
template <int OFF, int NR>
__device__ __forceinline__ float chunk_select(const uint32_t (&r)[NR], int j) {
    uint32_t a[16], b[8], c[4], d[2];
    const bool b0 = j & 1, b1 = j & 2, b2 = j & 4, b3 = j & 8, b4 = j & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = b0 ? r[OFF + 2 * i + 1] : r[OFF + 2 * i];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = b1 ? a[2 * i + 1] : a[2 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) c[i] = b2 ? b[2 * i + 1] : b[2 * i];
#pragma unroll
    for (int i = 0; i < 2; ++i) d[i] = b3 ? c[2 * i + 1] : c[2 * i];
    return fu(b4 ? d[1] : d[0]);
}

// Per-thread view of the row's filter_pairs_csr slice: a window of four consecutive viewed ids starting at index `cur`
// (B200_PAD_ID past the end of the slice).  Candidates arrive in ascending id order, so a window is only ever moved
// forward; its loads are issued when a candidate turns out to lie beyond it and are consumed one tile later
// (fifo_step), i.e. their latency is off the critical path.
struct CsrWindow {
    int64_t cur, fhi;
    int w0, w1, w2, w3;
    int streak;  // consecutive moves for the same candidate (long slices: switch to a binary search)
};

__device__ __forceinline__ void window_load(const int32_t* __restrict__ indices, CsrWindow& cw) {
    cw.w0 = cw.cur + 0 < cw.fhi ? __ldg(indices + cw.cur + 0) : B200_PAD_ID;
    cw.w1 = cw.cur + 1 < cw.fhi ? __ldg(indices + cw.cur + 1) : B200_PAD_ID;
    cw.w2 = cw.cur + 2 < cw.fhi ? __ldg(indices + cw.cur + 2) : B200_PAD_ID;
    cw.w3 = cw.cur + 3 < cw.fhi ? __ldg(indices + cw.cur + 3) : B200_PAD_ID;
}

// Move this thread's pending hits of chunk OFF (ascending column order) into its FIFO (a ring of T3_Q slots).  Returns
// true when some lane still has hits but no free slot: the caller runs a fifo_step and calls again with the remaining mask.
template <int OFF, int QN = T3_Q, int QS = T3_QSTRIDE, int NR>
__device__ __forceinline__ bool chunk_push(const uint32_t (&r)[NR], unsigned& hits, uint32_t pos0, float thr, uint32_t n_pos,
                                           uint32_t qaddr, int head, int& tail) {
    while (__any_sync(B200_FULL_MASK, hits != 0)) {
        if (hits && tail - head < QN) {
            const int j = __ffs(hits) - 1;
            hits &= hits - 1;
            // (taking the chunk maximum when it is the only score above the threshold, instead of the select tree, measured
            // slower: the extra branch costs more than the 31 selects it saves)
            const float val = chunk_select<OFF>(r, j);
            const uint32_t pos = pos0 + (uint32_t)(OFF + j);
            if (val > thr && pos < n_pos) {
                sts_v2(qaddr + (uint32_t)(tail & (QN - 1)) * QS, val, pos);
                ++tail;
            }
        }
        if (__any_sync(B200_FULL_MASK, hits != 0 && tail - head == QN)) return true;
    }
    return false;
}

// One step of the deferred work, for all 32 rows of the warp at once (no warp-collective inside: lanes may diverge):
// look at the oldest pending hit of the row; drop it if the threshold has passed it; if it lies beyond the CSR window,
// move the window (loads issued, not waited for) and leave the hit for the next step; otherwise test it against the
// window / the exclusion list and insert it into the row's candidate list.
template <int QN = T3_Q, int QS = T3_QSTRIDE>
__device__ __forceinline__ void fifo_step(const TcParams& p, RowState& rs, CsrWindow& cw, uint32_t qaddr, int& head, int tail,
                                          uint32_t ls, uint32_t li, int kc) {
    if (head == tail) return;
    float val;
    uint32_t pos;
    lds_v2(qaddr + (uint32_t)(head & (QN - 1)) * QS, val, pos);
    if (!(val > rs.thr)) {
        ++head;
        return;
    }
    const int obj = p.pos2obj ? __ldg(p.pos2obj + pos) : (int)pos;
    const int g = obj + p.id_off;
    if (g > cw.w3) {  // every id of the window is smaller (w3 == PAD_ID once the slice is exhausted: never taken then)
        cw.cur += 4;
        if (++cw.streak >= 2) {  // long slice: lower_bound of g in the rest
            int64_t lo = cw.cur, hi = cw.fhi;
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (__ldg(p.indices + mid) < g)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            cw.cur = lo;
        }
        window_load(p.indices, cw);
        return;
    }
    cw.streak = 0;
    ++head;
    const bool viewed = (g == cw.w0) | (g == cw.w1) | (g == cw.w2) | (g == cw.w3);
    if (!viewed && !(rs.xrow && is_excluded(rs, p.excl_n, g))) list_insert(ls, li, kc, rs, val, obj);
}

// Shared-memory map (dynamic, 1 KiB aligned): [KB] subject blocks | [NS] object blocks (16 KiB each: this CTA's half of a
// 256-object tile) | candidate lists [2 halves][128 rows][32] scores + ids | FIFOs | thresholds [2][128] | barriers.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T3_THREADS, 1)
tc3_topk_kernel(const __grid_constant__ CUtensorMap tm_sub, const __grid_constant__ CUtensorMap tm_obj, const TcParams p) {
    constexpr int BLKB_BYTES = T3_HALF * KBLK * 2;  // one object ring block: [128 rows][128 B]
    constexpr int NBUF = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int KB = p.kblocks, NS = p.n_stages;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)KB * BLK_BYTES;
    float* sLs = reinterpret_cast<float*>(sB + (size_t)NS * BLKB_BYTES);
    int* sLi = reinterpret_cast<int*>(sLs + 2 * TILE_M * 32);
    uint8_t* sQ = reinterpret_cast<uint8_t*>(sLi + 2 * TILE_M * 32);
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(sQ + T3_QBYTES);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sThr + 2 * TILE_M);
    const uint32_t bar_full = smem_u32(bars);
    const uint32_t bar_empty = smem_u32(bars + MAX_STAGES);
    const uint32_t bar_afull = smem_u32(bars + 2 * MAX_STAGES);
    const uint32_t bar_aempty = smem_u32(bars + 2 * MAX_STAGES + 1);
    const uint32_t bar_tfull = smem_u32(bars + 2 * MAX_STAGES + 2);
    const uint32_t bar_tempty = smem_u32(bars + 2 * MAX_STAGES + 2 + NBUF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 2 + 2 * NBUF);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader
    const int n_pairs = gridDim.x >> 1, pair = blockIdx.x >> 1;

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(bar_full + 8 * i, 1);   // leader's copy is the one that counts
            mbar_init(bar_empty + 8 * i, 1);  // one multicast commit per use
        }
        mbar_init(bar_afull, 1);
        mbar_init(bar_aempty, 1);
        for (int b = 0; b < NBUF; ++b) {
            mbar_init(bar_tfull + 8 * b, 1);
            mbar_init(bar_tempty + 8 * b, 16);  // 8 epilogue warps in each of the two CTAs arrive on the leader's copy
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_sub);
        tma_prefetch_desc(&tm_obj);
    }
    if (warp >= T3_EPI0) sts_thr(smem_u32(sThr + (warp - T3_EPI0) * 32 + lane), 0xffffffffu, INFINITY);  // tag no work item carries
    if (warp == 1) {
        tmem_alloc_2sm(smem_u32(tmem_slot), TMEM_COLS);
        tmem_relinquish_2sm();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tmem_base != 0) __trap();  // all 512 columns are ours

    const int n_work = p.n_row_tiles * p.n_splits;
    constexpr uint32_t BLK16 = BLK_BYTES >> 4;    // subject block in descriptor address units
    constexpr uint32_t BLKB16 = BLKB_BYTES >> 4;  // object block

    if (warp < T3_EPI0) reg_dealloc<T3_REGS_LOW>();  // all four warps of warp group 0
    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs, one elected thread)
        if (elect_one()) {
            uint32_t stage = 0, ph = 0, work_it = 0;
            const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
            for (int w = pair; w < n_work; w += n_pairs, ++work_it) {
                const int split = w / p.n_row_tiles, rt = w - split * p.n_row_tiles;
                const int t0 = split * p.tiles_per_split;
                const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
                if (work_it > 0) mbar_wait(bar_aempty, (work_it - 1) & 1);
                if (rank == 0) mbar_arrive_expect_tx(bar_afull, (uint32_t)(2 * KB * BLK_BYTES));
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d_2sm(sA_u + (uint32_t)kb * BLK_BYTES, &tm_sub, bar_afull, kb * KBLK, (rt * 2 + (int)rank) * TILE_M);
                const int nt = t1 - t0;
                const int ts = carousel_start(p, pair, work_it, split, t0, t1, rank == 0);
                for (int i = 0; i < nt; ++i) {
                    const int t = ts + i < t1 ? ts + i : ts + i - nt;
                    // the front is the position of the reference pair (pair 0 of each split's work items)
                    if (rank == 0 && p.front && (i & 15) == 0 && pair == 0)
                        *reinterpret_cast<volatile int32_t*>(p.front + split) = t;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * BLKB_BYTES);
                        tma_load_2d_2sm(sB_u + stage * BLKB_BYTES, &tm_obj, bar_full + 8 * stage, kb * KBLK,
                                        t * T3_TN + (int)rank * T3_HALF);
                        if (++stage == (uint32_t)NS) {
                            stage = 0;
                            ph ^= 1;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA only, one elected thread)
        if (rank == 0 && elect_one()) {
            uint32_t stage = 0, ph = 0, tile_it = 0, work_it = 0;
            const uint32_t a_lo0 = smem_desc_lo(smem_u32(sA)), b_lo0 = smem_desc_lo(smem_u32(sB));
            const uint32_t idesc = p.idesc;
            for (int w = pair; w < n_work; w += n_pairs, ++work_it) {
                const int split = w / p.n_row_tiles;
                const int t0 = split * p.tiles_per_split;
                const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
                mbar_wait(bar_afull, work_it & 1);
                tc_fence_after();
                for (int t = t0; t < t1; ++t, ++tile_it) {
                    const uint32_t buf = tile_it & 1, tph = (tile_it >> 1) & 1;
                    mbar_wait(bar_tempty + 8 * buf, tph ^ 1);  // both CTAs' epilogues have copied this accumulator out
                    tc_fence_after();
                    const uint32_t d0 = buf * (uint32_t)T3_TN;
                    uint32_t a_lo = a_lo0;
                    for (int kb = 0; kb < KB; ++kb, a_lo += BLK16) {
                        mbar_wait(bar_full + 8 * stage, ph);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + stage * BLKB16;
                        umma_f16_2sm(d0, a_lo, b_lo, SMEM_DESC_HI, idesc, (uint32_t)(kb != 0));
                        umma_f16_2sm(d0, a_lo + 2, b_lo + 2, SMEM_DESC_HI, idesc, 1u);
                        umma_f16_2sm(d0, a_lo + 4, b_lo + 4, SMEM_DESC_HI, idesc, 1u);
                        umma_f16_2sm(d0, a_lo + 6, b_lo + 6, SMEM_DESC_HI, idesc, 1u);
                        umma_commit_2sm(bar_empty + 8 * stage);  // frees this ring slot in both CTAs
                        if (++stage == (uint32_t)NS) {
                            stage = 0;
                            ph ^= 1;
                        }
                    }
                    umma_commit_2sm(bar_tfull + 8 * buf);
                }
                umma_commit_2sm(bar_aempty);
            }
        }
        __syncwarp();
    } else if (warp >= T3_EPI0) {
        // ===================================================================== epilogue (both CTAs): select candidates
        reg_alloc<T3_REGS_EPI>();
        const int ew = warp - T3_EPI0;
        const int half = ew >> 2, quarter = warp & 3;  // column half of the tile / TMEM lane quarter (== warp % 4)
        const int wrow0 = quarter * 32;                // first CTA-local subject row of this warp
        const uint32_t ls = pin(smem_u32(sLs + (size_t)(half * TILE_M + wrow0) * 32) + lane * 4);  // [slot][lane] arrays of this warp
        const uint32_t li = pin(smem_u32(sLi + (size_t)(half * TILE_M + wrow0) * 32) + lane * 4);
        const uint32_t qaddr = pin(smem_u32(sQ) + (uint32_t)(ew * 32 + lane) * 8);
        const uint32_t my_thr = pin(smem_u32(sThr + half * TILE_M + wrow0 + lane));
        const uint32_t peer_thr = pin(smem_u32(sThr + (half ^ 1) * TILE_M + wrow0 + lane));
        const uint32_t tempty0 = pin(mapa_rank(bar_tempty, 0)), tempty1 = pin(mapa_rank(bar_tempty + 8, 0));  // the leader's copies
        const uint32_t tfull0 = pin(bar_tfull), tfull1 = pin(bar_tfull + 8);
        const uint32_t tbase = pin(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * T3_HALF));
        const bool lane0 = pin((uint32_t)lane) == 0;
        const uint32_t n_pos = (uint32_t)p.n_pos;
        const int kc = p.k_cand;
        const bool dbg_skip = p.debug_mode == 2;
        uint32_t buf = 0, tph = 0, work_tag = 0;  // accumulator buffer / its phase parity: tile_it & 1, (tile_it >> 1) & 1
        for (int w = pair; w < n_work; w += n_pairs, ++work_tag) {
            const int split = w / p.n_row_tiles, rt = w - split * p.n_row_tiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
            const int64_t grow = ((int64_t)rt * 2 + rank) * TILE_M + wrow0 + lane;
            const bool row_ok = grow < p.n_rows;
            RowState rs;
            rs.thr = (row_ok && p.debug_mode == 0) ? -INFINITY : INFINITY;  // padded rows never produce candidates
            rs.cnt = 0;
            rs.minpos = 0;
            int head = 0, tail = 0;
            CsrWindow cw;
            sts_thr(my_thr, work_tag, rs.thr);
            const int nt = t1 - t0;
            int ts = 0;
            if (lane == 0) ts = carousel_start(p, pair, work_tag, split, t0, t1, false);
            ts = __shfl_sync(B200_FULL_MASK, ts, 0);
            const int64_t frow = row_ok ? (p.row_ids ? (int64_t)p.row_ids[grow] : grow) : -1;
            auto cursors_at = [&](int tile) {  // (re)position the CSR / exclusion cursors at the first object of `tile`
                const int64_t pos_first = (int64_t)tile * T3_TN + half * T3_HALF;
                const bool live = frow >= 0 && pos_first < p.n_pos;
                const int g_first = live ? (p.pos2obj ? __ldg(p.pos2obj + pos_first) : (int)pos_first) + p.id_off : 0;
                row_cursors_init(p, rs, live ? frow : -1, g_first);
                cw.cur = rs.cur;
                cw.fhi = rs.fhi;
                cw.streak = 0;
                window_load(p.indices, cw);
            };
            cursors_at(ts);
            int t = ts;
            uint32_t pos_t = (uint32_t)ts * T3_TN + (uint32_t)(half * T3_HALF);
            for (int it = 0; it < nt; ++it) {
                // exchange thresholds with the thread that owns the other column half of this row (monotone, racy by
                // design: a stale value is only a weaker bound; the tag keeps a value of the previous work item out)
                {
                    sts_thr(my_thr, work_tag, rs.thr);
                    uint32_t ptag;
                    float pthr;
                    lds_thr(peer_thr, ptag, pthr);
                    if (ptag == work_tag) rs.thr = fmaxf(rs.thr, pthr);
                }
                mbar_wait(buf ? tfull1 : tfull0, tph);
                tc_fence_after();
                // the last tile before the stream wraps around / of the work item: every pending hit must be handled
                // before the cursors are repositioned or the list is written
                const bool force = (t + 1 == t1);
                const bool last = (it + 1 == nt);
                if (!dbg_skip) {
                    uint32_t r[T3_HALF];
                    tmem_ld128_sync(tbase + buf * (uint32_t)T3_TN, r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive_cluster(buf ? tempty1 : tempty0);  // accumulator free again
                    const float m0 = chunk_max<0>(r), m1 = chunk_max<32>(r), m2 = chunk_max<64>(r), m3 = chunk_max<96>(r);
                    const float mx = fmaxf(max3(m0, m1, m2), m3);
                    const bool hit = __any_sync(B200_FULL_MASK, mx > rs.thr);
                    if (hit) {
                        const float thr = rs.thr;
                        unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0;
                        if (__any_sync(B200_FULL_MASK, m0 > thr)) h0 = chunk_hits<0>(r, thr);
                        if (__any_sync(B200_FULL_MASK, m1 > thr)) h1 = chunk_hits<32>(r, thr);
                        if (__any_sync(B200_FULL_MASK, m2 > thr)) h2 = chunk_hits<64>(r, thr);
                        if (__any_sync(B200_FULL_MASK, m3 > thr)) h3 = chunk_hits<96>(r, thr);
                        for (;;) {
                            bool stuck = chunk_push<0>(r, h0, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) stuck = chunk_push<32>(r, h1, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) stuck = chunk_push<64>(r, h2, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) stuck = chunk_push<96>(r, h3, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) break;
                            fifo_step(p, rs, cw, qaddr, head, tail, ls, li, kc);  // dense phase: make room, then go on
                        }
                    }
                    // deferred work: at most one step per tile (bounded latency in front of the next accumulator), except
                    // where everything pending has to be finished
#if B200_T3_STEP_LAZY
                    // (hits wait for at most STEP_PERIOD tiles; rows with a backlog and the flush points are served at once)
                    const bool due = (tail - head >= B200_T3_BACKLOG) ||
                                     (head != tail && ((it & (B200_T3_STEP_PERIOD - 1)) == B200_T3_STEP_PERIOD - 1 || force || last));
#else
                    const bool due = head != tail;
#endif
                    if (__any_sync(B200_FULL_MASK, due)) {
                        fifo_step(p, rs, cw, qaddr, head, tail, ls, li, kc);
                        if (force || last)
                            while (__any_sync(B200_FULL_MASK, head != tail)) fifo_step(p, rs, cw, qaddr, head, tail, ls, li, kc);
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive_cluster(buf ? tempty1 : tempty0);
                }
                buf ^= 1;
                tph ^= (buf == 0) ? 1u : 0u;
                ++t;
                pos_t += T3_TN;
                if (t == t1 && !last) {  // wrapped around: objects ascend again from the split's first tile
                    t = t0;
                    pos_t = (uint32_t)t0 * T3_TN + (uint32_t)(half * T3_HALF);
                    cursors_at(t0);
                }
            }
            // ---- write this thread's candidate list (unsorted): list index = split * 2 + column half
            if (row_ok) {
                const int64_t lrow = (int64_t)(split * 2 + half) * p.rows_pad + grow;
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
    cluster_sync_all();  // no CTA may exit (or free TMEM) while its peer can still signal its barriers / read its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace b200
