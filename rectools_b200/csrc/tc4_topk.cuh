// EXPERIMENTAL, NOT YET RUN ON HARDWARE (written after the round-1 GPU budget was spent; selected only by
// B200_TC_KERNEL=4, excluded from the tests and the bench default).  Same pipeline and epilogue protocol as tc3_topk.cuh,
// different geometry: SIXTEEN epilogue warps per CTA, each owning 32 subject rows x 64 accumulator columns, i.e. four
// candidate lists per row (16 slots each, K' <= 16) instead of two.
//
// Why: with 8 epilogue warps the kernel is bound by the *extraction* of hits -- ~110 dependent
// instructions at 5-6 cycles each with one runnable warp per scheduler -- on top of ~600 cycles of load + scan, against
// a tile time of 1480 cycles; every epilogue warp runs at ~75 % utilisation and a pair stalls whenever one of its 16
// warps is late.  Halving the slice halves load + scan per warp and, with K' = 8 per list, the hits per warp
// (4 x 8 / n per row instead of 2 x 12 / (n/2) ... = 32/n vs 48/n in total), so every warp runs below 50 % utilisation,
// where the pipeline model of section 7 predicts ~1520 cycles per tile at N = 1M (measured now: ~1990) and ~1800 at
// N_g = 125 K (now ~3470); four warps per scheduler also hide each other's latencies.
//
// Register budget: 20 warps = 5 per SM sub-partition (16384 registers): warp group 0 (TMA, MMA, 2 idle) 24 per thread,
// the four epilogue warp groups 120 per thread (64 staged scores + row state); launch at 96.
// Shared memory: lists [4 column quarters][128 rows][16 slots] x (score, id) = 64 KiB (as tc3), FIFOs 512 threads x
// T4_Q x 8 B, thresholds [4][128] x 8 B.
//
// The second-chance pass (K' = 32) cannot use 16-slot lists: the engine runs it on tc3_topk_kernel.
#pragma once
#include "tc3_topk.cuh"

namespace b200 {
namespace tc {

constexpr int T4_THREADS = 640;  // warp group 0: warps 0..3; epilogue: warps 4..19
constexpr int T4_EPI0 = 4;
constexpr int T4_EPI_WARPS = 16;
constexpr int T4_REGS_LOW = 24, T4_REGS_EPI = 120;  // 32 * (24 + 4 * 120) = 16128 <= 16384 per sub-partition
constexpr int T4_COLS = 64;                          // accumulator columns per epilogue thread and tile
constexpr int T4_SLOTS = 16;                         // list slots per (row, column quarter): K' <= 16
#ifndef B200_T4_Q
#define B200_T4_Q 4
#endif
constexpr int T4_Q = B200_T4_Q;
constexpr int T4_QSTRIDE = T4_EPI_WARPS * 32 * 8;    // bytes between FIFO slots: [slot][epilogue thread] x (score, position)
constexpr int T4_QBYTES = T4_Q * T4_QSTRIDE;
constexpr int T4_LIST_BYTES = 4 * TILE_M * T4_SLOTS * 4;  // one of the two arrays (scores / ids)
#ifndef B200_T4_STEP_PERIOD
#define B200_T4_STEP_PERIOD 16
#endif
#ifndef B200_T4_BACKLOG
#define B200_T4_BACKLOG 2
#endif

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T4_THREADS, 1)
tc4_topk_kernel(const __grid_constant__ CUtensorMap tm_sub, const __grid_constant__ CUtensorMap tm_obj, const TcParams p) {
    constexpr int BLKB_BYTES = T3_HALF * KBLK * 2;  // one object ring block: [128 rows][128 B] (this CTA's half of a tile)
    constexpr int NBUF = 2;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int KB = p.kblocks, NS = p.n_stages;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)KB * BLK_BYTES;
    float* sLs = reinterpret_cast<float*>(sB + (size_t)NS * BLKB_BYTES);  // [4][128 rows][16 slots], per warp [slot][lane]
    int* sLi = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(sLs) + T4_LIST_BYTES);
    uint8_t* sQ = reinterpret_cast<uint8_t*>(sLi) + T4_LIST_BYTES;
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(sQ + T4_QBYTES);  // [4 column quarters][128 rows]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sThr + 4 * TILE_M);
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
            mbar_init(bar_tempty + 8 * b, 2 * T4_EPI_WARPS);  // 16 epilogue warps in each of the two CTAs arrive on the leader's copy
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_sub);
        tma_prefetch_desc(&tm_obj);
    }
    if (warp >= T4_EPI0) sts_thr(smem_u32(sThr + (warp - T4_EPI0) * 32 + lane), 0xffffffffu, INFINITY);  // tag no work item carries
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

    if (warp < T4_EPI0) reg_dealloc<T4_REGS_LOW>();  // all four warps of warp group 0
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
    } else if (warp >= T4_EPI0) {
        // ===================================================================== epilogue (both CTAs): select candidates
        reg_alloc<T4_REGS_EPI>();
        const int ew = warp - T4_EPI0;                 // 0..15
        const int colq = ew >> 2, quarter = warp & 3;  // column quarter of the tile / TMEM lane quarter (== warp % 4)
        const int wrow0 = quarter * 32;                // first CTA-local subject row of this warp
        // [slot][lane] arrays of this warp: 16 slots x 32 lanes x 4 B = 2 KiB, slot stride 128 B (as list_insert expects)
        const uint32_t ls = pin(smem_u32(sLs) + (uint32_t)((colq * TILE_M + wrow0) * T4_SLOTS * 4) + lane * 4);
        const uint32_t li = pin(smem_u32(sLi) + (uint32_t)((colq * TILE_M + wrow0) * T4_SLOTS * 4) + lane * 4);
        const uint32_t qaddr = pin(smem_u32(sQ) + (uint32_t)(ew * 32 + lane) * 8);
        const uint32_t thr_row = pin(smem_u32(sThr + wrow0 + lane));  // + colq * 128 * 8: the four threads of this row
        const uint32_t my_thr = pin(thr_row + (uint32_t)colq * (TILE_M * 8));
        const uint32_t tempty0 = pin(mapa_rank(bar_tempty, 0)), tempty1 = pin(mapa_rank(bar_tempty + 8, 0));  // the leader's copies
        const uint32_t tfull0 = pin(bar_tfull), tfull1 = pin(bar_tfull + 8);
        const uint32_t tbase = pin(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colq * T4_COLS));
        const bool lane0 = pin((uint32_t)lane) == 0;
        const uint32_t n_pos = (uint32_t)p.n_pos;
        const int kc = min(p.k_cand, T4_SLOTS);
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
                const int64_t pos_first = (int64_t)tile * T3_TN + colq * T4_COLS;
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
            uint32_t pos_t = (uint32_t)ts * T3_TN + (uint32_t)(colq * T4_COLS);
            for (int it = 0; it < nt; ++it) {
                // exchange thresholds with the three threads that own the other column quarters of this row (monotone,
                // racy by design: a stale value is only a weaker bound; the tag keeps a value of the previous work item out)
                {
                    sts_thr(my_thr, work_tag, rs.thr);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t ptag;
                        float pthr;
                        lds_thr(thr_row + (uint32_t)q * (TILE_M * 8), ptag, pthr);  // (q == colq reads back the own value)
                        if (ptag == work_tag) rs.thr = fmaxf(rs.thr, pthr);
                    }
                }
                mbar_wait(buf ? tfull1 : tfull0, tph);
                tc_fence_after();
                const bool force = (t + 1 == t1);
                const bool last = (it + 1 == nt);
                if (!dbg_skip) {
                    uint32_t r[T4_COLS];
                    tmem_ld64_sync(tbase + buf * (uint32_t)T3_TN, r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive_cluster(buf ? tempty1 : tempty0);  // accumulator free again
                    const float m0 = chunk_max<0>(r), m1 = chunk_max<32>(r);
                    const float mx = fmaxf(m0, m1);
                    const bool hit = __any_sync(B200_FULL_MASK, mx > rs.thr);
                    if (hit) {
                        const float thr = rs.thr;
                        unsigned h0 = 0, h1 = 0;
                        if (__any_sync(B200_FULL_MASK, m0 > thr)) h0 = chunk_hits<0>(r, thr);
                        if (__any_sync(B200_FULL_MASK, m1 > thr)) h1 = chunk_hits<32>(r, thr);
                        for (;;) {
                            bool stuck = chunk_push<0, T4_Q, T4_QSTRIDE>(r, h0, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) stuck = chunk_push<32, T4_Q, T4_QSTRIDE>(r, h1, pos_t, rs.thr, n_pos, qaddr, head, tail);
                            if (!stuck) break;
                            fifo_step<T4_Q, T4_QSTRIDE>(p, rs, cw, qaddr, head, tail, ls, li, kc);  // dense phase: make room, then go on
                        }
                    }
                    const bool due = (tail - head >= B200_T4_BACKLOG) ||
                                     (head != tail && ((it & (B200_T4_STEP_PERIOD - 1)) == B200_T4_STEP_PERIOD - 1 || force || last));
                    if (__any_sync(B200_FULL_MASK, due)) {
                        fifo_step<T4_Q, T4_QSTRIDE>(p, rs, cw, qaddr, head, tail, ls, li, kc);
                        if (force || last)
                            while (__any_sync(B200_FULL_MASK, head != tail))
                                fifo_step<T4_Q, T4_QSTRIDE>(p, rs, cw, qaddr, head, tail, ls, li, kc);
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
                    pos_t = (uint32_t)t0 * T3_TN + (uint32_t)(colq * T4_COLS);
                    cursors_at(t0);
                }
            }
            // ---- write this thread's candidate list (unsorted): list index = split * 4 + column quarter; the global layout
            // keeps 32 slots per list (select_kernel), the upper 16 are padding
            if (row_ok) {
                const int64_t lrow = (int64_t)(split * 4 + colq) * p.rows_pad + grow;
                for (int e = 0; e < T4_SLOTS; ++e) {
                    const bool keep = e < rs.cnt;
                    float sv = -INFINITY;
                    int iv = B200_PAD_ID;
                    if (keep) {
                        sv = lds_f32(ls + e * 128);
                        iv = lds_s32(li + e * 128);
                    }
                    p.cand_scores[lrow * 32 + e] = sv;
                    p.cand_ids[lrow * 32 + e] = iv;
                }
                for (int e = T4_SLOTS; e < 32; ++e) {
                    p.cand_scores[lrow * 32 + e] = -INFINITY;
                    p.cand_ids[lrow * 32 + e] = B200_PAD_ID;
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
