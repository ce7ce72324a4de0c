// 2-SM variant of the fused scoring + candidate-selection kernel: CTA pairs (cluster of 2) with `tcgen05.mma.cta_group::2`.
//
// Why: the 1-SM kernel (tc_topk.cuh) issues 128x128x16 MMAs whose two shared-memory operands cost 8 KiB per 64-cycle
// instruction = 128 B/clk, the whole shared-memory bandwidth of an SM; measured MMA-only rate 1.27 PFLOP/s (76 % of the
// measured cuBLAS peak) with the epilogue disabled.  A CTA pair computes a [256 subjects x 256 objects] tile per
// 256x256x16 MMA: each CTA keeps its own 128 subject rows resident, loads only HALF of every object tile (the tensor
// cores read the other half from the peer's shared memory) and owns the [128 x 256] fp32 accumulator rows of its
// subjects in its own TMEM (2 buffers x 256 columns).  Per CTA that is 8 KiB of operands per 128-cycle instruction
// (64 B/clk) and half the L2->SM traffic per FLOP.
//
// Roles per CTA (384 threads): warp 0 TMA producer (own subject rows + own half of the object tiles, completion
// counted on the LEADER's mbarriers), warp 1 MMA issuer (leader CTA only; commits are multicast to both CTAs),
// warp 2 TMEM allocator, warps 4..11 epilogue.  Epilogue warp e owns TMEM lanes 32*(e%4).. and the column half e/4 of
// every tile, i.e. each (row, column-half) has its own candidate list and running threshold; the two thresholds of a
// row are exchanged through shared memory so that either list prunes with the better of the two.
//
// Replaces the same reference code as tc_topk.cuh (rank_implicit.py:264-272 / rank_torch.py:133-152) as a candidate
// generator; final scores / order / certificate come from select_kernel<true>.
#pragma once
#include "tc_topk.cuh"

namespace b200 {
namespace tc {


__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Arrive on the mbarrier at the same shared-memory offset in CTA `rank` of the cluster.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank)
        : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's shared memory, the byte count is credited to the LEADER CTA's mbarrier
// (shared::cta addresses carry the CTA-pair rank in bit 24; clearing it names the even CTA's copy of the barrier).
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[both CTAs' smem, 128 rows each] * B[both CTAs' smem, 128 rows each]^T : 256 x 256 x 16.
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                             uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accum)
        : "memory");
}
// Arrive (once the MMAs issued so far have retired) on the mbarrier at this offset in BOTH CTAs of the pair.
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

// Start tile of a work item: decided once by the leader CTA's producer thread (the current front of its object split),
// published through global memory, read by every other role of both CTAs.
__device__ __forceinline__ int carousel_start(const TcParams& p, int pair, uint32_t work_it, int split, int t0, int t1, bool decide) {
    if (p.front == nullptr) return t0;
    volatile int32_t* slot = p.starts + (size_t)pair * p.starts_stride + work_it;
    if (decide) {
        int s = *reinterpret_cast<volatile int32_t*>(p.front + split);
        s = min(max(s, t0), t1 - 1);
        *slot = s;
        __threadfence();
        return s;
    }
    int s;
    for (uint32_t spins = 0; (s = *slot) < 0; ++spins)
        if (spins > (1u << 26)) __trap();
    return s;
}

// TN   = objects per tile of the CTA pair (each CTA loads TN/2 of them and owns TN accumulator columns per buffer)
// NBUF = accumulator buffers in TMEM (NBUF * TN = 512 columns): 2 x 256 minimises MMA instructions and shared-memory
//        operand traffic (64 B/clk per CTA), 4 x 128 (96 B/clk) lets the MMA run up to three tiles ahead of a warp
//        that is busy inserting candidates.
// STAGE = true: an epilogue warp first copies its whole [32 rows x TN/2 columns] slice of the accumulator to registers
//        (one wide tcgen05.ld), hands the TMEM buffer back at once and only then scans the scores, so the MMA of tile
//        t+2 never waits for candidate insertion of tile t (352 threads per CTA to afford ~180 registers per thread).
template <bool STAGE>
struct Tc2Threads {
    static constexpr int EPI0 = STAGE ? 3 : EPI_WARP0;  // first epilogue warp; (warp & 3) is its TMEM lane quarter
    static constexpr int THREADS = (EPI0 + 8) * 32;
};

template <int TN, int NBUF, bool STAGE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(Tc2Threads<STAGE>::THREADS, 1)
tc2_topk_kernel(const __grid_constant__ CUtensorMap tm_sub, const __grid_constant__ CUtensorMap tm_obj, const TcParams p) {
    static_assert(TN * NBUF == TMEM_COLS && (TN == 256 || TN == 128), "accumulators must fill the 512 TMEM columns");
    constexpr int TILE2_N = TN;
    constexpr int HALF_N = TN / 2;                 // object rows per CTA and tile = accumulator columns per epilogue warp
    constexpr int BLKB_BYTES = HALF_N * KBLK * 2;  // one object ring block: [TN/2 rows][128 B]
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int KB = p.kblocks, NS = p.n_stages;
    uint8_t* sA = smem;                                      // [KB] blocks: this CTA's 128 subject rows
    uint8_t* sB = sA + (size_t)KB * BLK_BYTES;               // [NS] blocks: this CTA's half of the object tiles
    float* sLs = reinterpret_cast<float*>(sB + (size_t)NS * BLKB_BYTES);  // [2 halves][128 rows][32] candidate scores
    int* sLi = reinterpret_cast<int*>(sLs + 2 * TILE_M * 32);            // [2][128][32] candidate ids
    // [2][128] published (threshold, work-item tag) pairs: the tag keeps a warp that has already moved on to the next
    // subject tile from adopting its partner's threshold of the previous one
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(sLi + 2 * TILE_M * 32);
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
    constexpr int EPI0 = Tc2Threads<STAGE>::EPI0;
    if (warp >= EPI0) sThr[(warp - EPI0) * 32 + lane] = ~0ull;  // tag no work item can carry
    if (warp == 2) {
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
                    // (measured: letting every pair overwrite it does not re-align pairs that drifted apart -- 333 GB of DRAM
                    // reads per launch at U = 1M instead of 15 GB)
                    if (rank == 0 && p.front && (i & 15) == 0 && pair == 0)
                        *reinterpret_cast<volatile int32_t*>(p.front + split) = t;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * BLKB_BYTES);
                        tma_load_2d_2sm(sB_u + stage * BLKB_BYTES, &tm_obj, bar_full + 8 * stage, kb * KBLK,
                                        t * TILE2_N + (int)rank * HALF_N);
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
        // ===================================================================== MMA issuer (leader CTA only)
        // One elected thread runs the whole role (waits included): measured with per-k-block election the issue loop cost
        // ~315 cycles per 4 MMAs (256 cycles of tensor work) and the tensor pipe sat at 45 %.
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
                    const uint32_t buf = tile_it % NBUF, tph = (tile_it / NBUF) & 1;
                    mbar_wait(bar_tempty + 8 * buf, tph ^ 1);  // both CTAs' epilogues have drained this accumulator
                    tc_fence_after();
                    const uint32_t d0 = buf * (uint32_t)TILE2_N;
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
    } else if (warp >= EPI0) {
        // ===================================================================== epilogue (both CTAs): select candidates
        const int ew = warp - EPI0;
        const int half = ew >> 2, quarter = warp & 3;  // column half of the tile / TMEM lane quarter (== warp % 4)
        const int wrow0 = quarter * 32;              // first CTA-local subject row of this warp
        const uint32_t ls = smem_u32(sLs + (size_t)(half * TILE_M + wrow0) * 32) + lane * 4;  // [slot][lane] arrays of this warp
        const uint32_t li = smem_u32(sLi + (size_t)(half * TILE_M + wrow0) * 32) + lane * 4;
        volatile unsigned long long* myThr = sThr + half * TILE_M + wrow0 + lane;
        const volatile unsigned long long* peerThr = sThr + (half ^ 1) * TILE_M + wrow0 + lane;
        uint32_t work_tag = 0;
        const int kc = p.k_cand;
        uint32_t tile_it = 0;
        for (int w = pair; w < n_work; w += n_pairs, ++work_tag) {
            const int split = w / p.n_row_tiles, rt = w - split * p.n_row_tiles;
            const int t0 = split * p.tiles_per_split;
            const int t1 = min(t0 + p.tiles_per_split, p.n_obj_tiles);
            const int64_t grow0 = ((int64_t)rt * 2 + rank) * TILE_M + wrow0;  // global row of lane 0
            const int64_t grow = grow0 + lane;
            const bool row_ok = grow < p.n_rows;
            RowState rs;
            rs.thr = (row_ok && p.debug_mode == 0) ? -INFINITY : INFINITY;
            rs.cnt = 0;
            rs.minpos = 0;
            *myThr = ((unsigned long long)work_tag << 32) | __float_as_uint(rs.thr);
            __syncwarp();
            const int nt = t1 - t0;
            int ts = 0;
            if (lane == 0) ts = carousel_start(p, pair, work_tag, split, t0, t1, false);
            ts = __shfl_sync(B200_FULL_MASK, ts, 0);
            const int64_t frow = row_ok ? (p.row_ids ? (int64_t)p.row_ids[grow] : grow) : -1;
            auto cursors_at = [&](int tile) {  // (re)position the CSR / exclusion cursors at the first object of `tile`
                const int64_t pos_first = (int64_t)tile * TILE2_N + half * HALF_N;
                const bool live = frow >= 0 && pos_first < p.n_pos;
                const int g_first = live ? (p.pos2obj ? __ldg(p.pos2obj + pos_first) : (int)pos_first) + p.id_off : 0;
                row_cursors_init(p, rs, live ? frow : -1, g_first);
            };
            cursors_at(ts);
            for (int it = 0; it < nt; ++it, ++tile_it) {
                const int t = ts + it < t1 ? ts + it : ts + it - nt;
                if (it > 0 && t == t0) cursors_at(t0);  // wrapped around: objects ascend again from the split's first tile
                const uint32_t buf = tile_it % NBUF, tph = (tile_it / NBUF) & 1;
                // exchange thresholds with the thread that owns the other column half of this row (monotone, racy by
                // design: a stale value is only a weaker bound)
                *myThr = ((unsigned long long)work_tag << 32) | __float_as_uint(rs.thr);
                {
                    const unsigned long long pv = *peerThr;
                    if ((uint32_t)(pv >> 32) == work_tag) rs.thr = fmaxf(rs.thr, __uint_as_float((uint32_t)pv));
                }
                mbar_wait(bar_tfull + 8 * buf, tph);
                tc_fence_after();
                const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(buf * TILE2_N + half * HALF_N);
                const int64_t pos_t = (int64_t)t * TILE2_N + half * HALF_N;
                if (p.debug_mode == 2) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
                    continue;
                }
                if constexpr (STAGE) {
                    uint32_t r[HALF_N];
                    if constexpr (HALF_N == 128)
                        tmem_ld128_sync(tbase, r);
                    else
                        tmem_ld64_sync(tbase, r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);  // accumulator free again
                    process_chunk<0, HALF_N>(r, pos_t, p, ls, li, kc, rs);
                    process_chunk<32, HALF_N>(r, pos_t + 32, p, ls, li, kc, rs);
                    if constexpr (HALF_N == 128) {
                        process_chunk<64, HALF_N>(r, pos_t + 64, p, ls, li, kc, rs);
                        process_chunk<96, HALF_N>(r, pos_t + 96, p, ls, li, kc, rs);
                    }
                } else {
                    uint32_t ra[32], rb[32];
                    tmem_ld_issue(tbase, ra);
#pragma unroll 1
                    for (int h = 0; h < HALF_N / 64; ++h) {
                        tmem_ld_wait(ra);
                        tmem_ld_issue(tbase + h * 64 + 32, rb);
                        process_chunk(ra, pos_t + h * 64, p, ls, li, kc, rs);
                        tmem_ld_wait(rb);
                        if (h + 1 < HALF_N / 64) {
                            tmem_ld_issue(tbase + (h + 1) * 64, ra);
                        } else {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_remote(bar_tempty + 8 * buf, 0);
                        }
                        process_chunk(rb, pos_t + h * 64 + 32, p, ls, li, kc, rs);
                    }
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
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, TMEM_COLS);
    }
}

}  // namespace tc
}  // namespace b200
