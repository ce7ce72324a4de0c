// The fused scoring + candidate-selection kernel: one persistent CTA pair per two SMs.
//
//   TMA (own 128 subject rows once per work item; own half of every 256-object tile through a ring of 16 KiB blocks)
//   -> tcgen05.mma.cta_group::2 256 x 256 x 16 (fp16 / bf16 -> fp32), two 256-column accumulators per CTA in TMEM
//   -> tcgen05.ld of a [32 rows x COLS columns] slice per epilogue warp into registers, accumulator handed back at once
//   -> threshold scan (3-input max tree, ONE vote per tile), hits extracted into per-thread ring FIFOs in shared memory
//   -> deferred, bounded steps: filter_pairs_csr lookup through a prefetched 4-entry window, candidate-list insertion.
// Score rows never reach HBM: only the K' best (score, id) pairs per row and column group are written.
//
// NW = epilogue warps per CTA:
//   8  : 32 rows x 128 columns per warp, two candidate lists per row (32 slots each), 232 registers per epilogue thread
//   16 : 32 rows x  64 columns per warp, four lists per row (16 slots each), 120 registers; four warps per scheduler
//        hide each other's latencies and a warp is hit by half as many candidates.
// Warp group 0: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator (leader CTA issues), warps 2-3 = peer-threshold
// helpers (multi-GPU: they poll the other ranks' published thresholds over NVLink and publish this rank's; idle otherwise).
// `setmaxnreg` moves the registers warp group 0 does not need to the epilogue warps.
//
// Three selection modes share the epilogue:
//   * adaptive lists (k <= 24): replace-minimum lists of K' slots, the list minimum is the running threshold;
//   * wide mode (24 < k <= 128): adaptive lists for the first `phase1_tiles` tiles of the stream, then the threshold is
//     FROZEN and every later score above it is appended to a global list -- ~1.5 k candidates per row in ONE pass with
//     ~3x fewer hits than adaptive lists of that size would take;
//   * shared thresholds (item-sharded multi-GPU): the row's threshold is the maximum over all ranks' thresholds.
// In every mode a list's final threshold bounds every score the list ever discarded: the certificate of select.cuh.
//
// Replaces the same reference code as named in tc_common.cuh (rank_implicit.py:264-272 / rank_torch.py:133-152).
#pragma once
#include "tc_common.cuh"

namespace b200 {
namespace tc {

constexpr uint32_t TAG_NONE = 0xffffffffu, TAG_DONE = 0xfffffffeu;  // exchange-slot tags no work item carries

template <int NW>
struct FusedCfg {
    static_assert(NW == 8 || NW == 16, "8 or 16 epilogue warps");
    static constexpr int EPI0 = 4;                    // first epilogue warp; (warp & 3) is its TMEM lane quarter
    static constexpr int THREADS = (EPI0 + NW) * 32;
    static constexpr int COLS = 1024 / NW;            // accumulator columns per epilogue thread and tile
    static constexpr int NLIST = NW / 4;              // column groups = candidate lists per row
    static constexpr int SLOTS = 64 / NLIST;          // list capacity (K' <= SLOTS)
    // setmaxnreg moves registers inside the CTA's LAUNCH allocation (168 x 384 = 64512, 96 x 640 = 61440), not the whole file:
    // 4 * REGS_LOW + NW * REGS_EPI <= launch registers * (4 + NW).  (120 for the 16-warp geometry over-subscribes the pool: the
    // last epilogue warps wait for ever in setmaxnreg.inc -- the first hardware run of round 2.)
    static constexpr int REGS_LOW = NW == 8 ? 40 : 32;
    static constexpr int REGS_EPI = NW == 8 ? 232 : 112;
    static constexpr int Q = NW == 8 ? 8 : 4;         // deferred hits per thread (ring FIFO), measured with the step period
    static constexpr int QSTRIDE = NW * 32 * 8;       // bytes between FIFO slots: [slot][epilogue thread] x (score, position)
    static constexpr int QBYTES = Q * QSTRIDE;
    static constexpr int BACKLOG = NW == 8 ? 4 : 2;   // a row with this many pending hits gets a step at once
    static constexpr int PERIOD = 16;                 // otherwise deferred work runs every PERIOD-th tile (power of two)
    static constexpr int LIST_BYTES = NLIST * TILE_M * SLOTS * 4;  // one of the two arrays (scores / ids): 32 KiB
    static constexpr int THR_BYTES = (NLIST + 1) * TILE_M * 8;     // (tag, threshold) per list + one slot for the peers' maximum
    static constexpr int FIXED_BYTES = 2 * LIST_BYTES + QBYTES + THR_BYTES + 1024 /*alignment slack*/ + 512 /*barriers*/;
};

// Move this thread's pending hits of chunk OFF (ascending column order) into its FIFO (a ring of QN slots).  Returns
// true when some lane still has hits but no free slot: the caller runs a fifo_step and calls again with the remaining mask.
template <int OFF, int QN, int QS, int NR>
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

// Where accepted candidates go: the row's replace-minimum list in shared memory (adaptive) or, once the threshold is
// frozen, the next free slot of its global list.
struct Sink {
    float* gs;      // this thread's global list (scores / ids), `cap` slots
    int32_t* gi;
    int cap;
    bool appending;
};

// One step of the deferred work, for all 32 rows of the warp at once (no warp-collective inside: lanes may diverge):
// look at the oldest pending hit of the row; drop it if the threshold has passed it; if it lies beyond the CSR window,
// move the window (loads issued, not waited for) and leave the hit for the next step; otherwise test it against the
// window / the exclusion list and hand it to the sink.
template <int QN, int QS, bool WIDE>
__device__ __forceinline__ void fifo_step(const TcParams& p, RowState& rs, CsrWindow& cw, uint32_t qaddr, int& head, int tail,
                                          uint32_t ls, uint32_t li, int kc, const Sink& sink) {
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
    if (!viewed && !(rs.xrow && is_excluded(rs, p.excl_n, g))) {
        if (WIDE && sink.appending) {
            if (rs.cnt < sink.cap) {
                sink.gs[rs.cnt] = val;
                sink.gi[rs.cnt] = obj;
            }
            ++rs.cnt;  // beyond the capacity: counted, not stored -- the row fails its certificate and is re-ranked
        } else {
            list_insert(ls, li, kc, rs, val, obj);
        }
    }
}

// Shared-memory map (dynamic, 1 KiB aligned): [KB] subject blocks | [NS] object blocks (16 KiB each: this CTA's half of a
// 256-object tile) | candidate lists [NLIST][128 rows][SLOTS] scores + ids | FIFOs | thresholds [NLIST + 1][128] | barriers.
// WIDE / PEERS compile the wide mode (threshold freeze + global append) and the peer-threshold exchange in; the plain
// instantiation carries neither in its tile loop (measured: the run-time switches cost the 8-warp kernel ~6 %).
template <int NW, bool WIDE, bool PEERS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FusedCfg<NW>::THREADS, 1)
fused_topk_kernel(const __grid_constant__ CUtensorMap tm_sub, const __grid_constant__ CUtensorMap tm_obj, const TcParams p) {
    using Cfg = FusedCfg<NW>;
    constexpr int NBUF = 2;
    constexpr int COLS = Cfg::COLS, NLIST = Cfg::NLIST, SLOTS = Cfg::SLOTS, QN = Cfg::Q, QS = Cfg::QSTRIDE;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);

    const int KB = p.kblocks, NS = p.n_stages;
    uint8_t* sA = smem;
    uint8_t* sB = sA + (size_t)KB * BLK_BYTES;
    uint8_t* sLs = sB + (size_t)NS * BLK_BYTES;  // per warp [slot][lane] arrays
    uint8_t* sLi = sLs + Cfg::LIST_BYTES;
    uint8_t* sQ = sLi + Cfg::LIST_BYTES;
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(sQ + Cfg::QBYTES);  // [NLIST + 1][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sThr + (NLIST + 1) * TILE_M);
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
            mbar_init(bar_tempty + 8 * b, 2 * NW);  // the epilogue warps of both CTAs arrive on the leader's copy
        }
        fence_barrier_init();
        tma_prefetch_desc(&tm_sub);
        tma_prefetch_desc(&tm_obj);
    }
    for (int i = threadIdx.x; i < (NLIST + 1) * TILE_M; i += blockDim.x) sts_thr(smem_u32(sThr + i), TAG_NONE, INFINITY);
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
    constexpr uint32_t BLK16 = BLK_BYTES >> 4;  // a 16 KiB block in descriptor address units

    if (warp < Cfg::EPI0) reg_dealloc<Cfg::REGS_LOW>();  // all four warps of warp group 0
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
                    // the front is the position of the reference pair (pair 0 of each split's work items); measured: letting
                    // every pair overwrite it does not re-align pairs that drifted apart (333 GB of DRAM reads instead of 15)
                    if (rank == 0 && p.front && (i & 15) == 0 && pair == 0)
                        *reinterpret_cast<volatile int32_t*>(p.front + split) = t;
                    for (int kb = 0; kb < KB; ++kb) {
                        mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * BLK_BYTES);
                        tma_load_2d_2sm(sB_u + stage * BLK_BYTES, &tm_obj, bar_full + 8 * stage, kb * KBLK,
                                        t * TILE_N + (int)rank * HALF_N);
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
        // One thread runs the whole role, waits included: with per-k-block election the issue loop cost ~315 cycles per
        // 4 MMAs (256 cycles of tensor work) and the tensor pipe sat at 45 %.
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
                    const uint32_t d0 = buf * (uint32_t)TILE_N;
                    uint32_t a_lo = a_lo0;
                    for (int kb = 0; kb < KB; ++kb, a_lo += BLK16) {
                        mbar_wait(bar_full + 8 * stage, ph);
                        tc_fence_after();
                        const uint32_t b_lo = b_lo0 + stage * BLK16;
                        // +32 B per K = 16 step inside the 128 B swizzle atom = +2 in descriptor address units
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
    } else if (warp < Cfg::EPI0) {
        // ===================================================================== peer-threshold helpers (warps 2 and 3)
        // Thread h serves CTA-local rows h and h + 64.  Per round and row: read the row's own thresholds from the exchange
        // slots, publish their maximum to this rank's global array, read the other ranks' published values (NVLink peer
        // loads, latency irrelevant here), leave their maximum in the row's extra exchange slot.  All values are monotone
        // lower bounds of the row's final threshold: a stale one is only weaker, never wrong.
        if (PEERS && p.n_peers > 0) {
            const int h = (warp - 2) * 32 + lane;
            float published[2] = {-INFINITY, -INFINITY};
            uint32_t pub_tag[2] = {0xffffffffu, 0xffffffffu};
            bool done[2] = {false, false};
            while (!(done[0] && done[1])) {
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const int r = h + 64 * s;
                    uint32_t tag;
                    float own;
                    lds_thr(smem_u32(sThr + r), tag, own);
                    done[s] = tag == TAG_DONE;  // the row's first column group has finished its last work item
                    if (tag >= TAG_DONE) continue;
#pragma unroll
                    for (int l = 1; l < NLIST; ++l) {
                        uint32_t tl;
                        float vl;
                        lds_thr(smem_u32(sThr + l * TILE_M + r), tl, vl);
                        if (tl == tag) own = fmaxf(own, vl);
                    }
                    const int w = pair + (int)tag * n_pairs;
                    if (w >= n_work) continue;
                    const int rt = w % p.n_row_tiles;
                    const int64_t grow = ((int64_t)rt * 2 + rank) * TILE_M + r;
                    if (grow >= p.n_rows) continue;
                    if (pub_tag[s] != tag) {
                        pub_tag[s] = tag;
                        published[s] = -INFINITY;
                    }
                    if (own > published[s] && own < INFINITY) {
                        published[s] = own;
                        stg_peer(p.peer_pub + p.peer_row0 + grow, p.peer_epoch, ldexpf(own, -p.peer_exp));
                    }
                    // all peer loads in flight at once: one NVLink round trip per row and round (issued one after the other the
                    // seven loads took ~15 us -- 20 tiles -- and the shared thresholds arrived after the warm-up they are meant
                    // to shorten: first 8-GPU measurement of round 2)
                    unsigned long long v[MAX_PEERS];
#pragma unroll
                    for (int q = 0; q < MAX_PEERS; ++q) v[q] = q < p.n_peers ? ldg_peer(p.peer_in[q] + p.peer_row0 + grow) : 0ull;
                    float best = -INFINITY;
#pragma unroll
                    for (int q = 0; q < MAX_PEERS; ++q)
                        if (q < p.n_peers && (uint32_t)(v[q] >> 32) == p.peer_epoch) best = fmaxf(best, __uint_as_float((uint32_t)v[q]));
                    if (best > -INFINITY) sts_thr(smem_u32(sThr + NLIST * TILE_M + r), tag, ldexpf(best, p.peer_exp));
                }
                __nanosleep(200);
            }
        }
    } else if (warp >= Cfg::EPI0) {
        // ===================================================================== epilogue (both CTAs): select candidates
        reg_alloc<Cfg::REGS_EPI>();
        const int ew = warp - Cfg::EPI0;
        const int colg = ew >> 2, quarter = warp & 3;  // column group of the tile / TMEM lane quarter (== warp % 4)
        const int wrow0 = quarter * 32;                // first CTA-local subject row of this warp
        // [slot][lane] arrays of this warp: SLOTS x 32 lanes x 4 B, slot stride 128 B (as list_insert expects)
        const uint32_t ls = pin(smem_u32(sLs) + (uint32_t)((colg * TILE_M + wrow0) * SLOTS * 4) + lane * 4);
        const uint32_t li = pin(smem_u32(sLi) + (uint32_t)((colg * TILE_M + wrow0) * SLOTS * 4) + lane * 4);
        const uint32_t qaddr = pin(smem_u32(sQ) + (uint32_t)(ew * 32 + lane) * 8);
        const uint32_t thr_row = pin(smem_u32(sThr + wrow0 + lane));  // + l * 128 * 8: the threads of this row, [NLIST]: peers
        const uint32_t my_thr = pin(thr_row + (uint32_t)colg * (TILE_M * 8));
        const uint32_t tempty0 = pin(mapa_rank(bar_tempty, 0)), tempty1 = pin(mapa_rank(bar_tempty + 8, 0));  // the leader's copies
        const uint32_t tfull0 = pin(bar_tfull), tfull1 = pin(bar_tfull + 8);
        const uint32_t tbase = pin(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(colg * COLS));
        const bool lane0 = pin((uint32_t)lane) == 0;
        const uint32_t n_pos = (uint32_t)p.n_pos;
        const int kc = p.k_cand;  // (<= SLOTS, guaranteed by the host; a visible bound makes the compiler unroll the list scans fully and spill)
        const bool dbg_skip = p.debug_mode == 2;
        const bool peers = PEERS && p.n_peers > 0;
        const uint32_t other_thr = pin(thr_row + (uint32_t)(colg ^ 1) * (TILE_M * 8));  // two lists per row: the other one's slot
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
            const int64_t lrow = (int64_t)(split * NLIST + colg) * p.rows_pad + (row_ok ? grow : 0);  // this thread's global list
            Sink sink;
            sink.gs = p.cand_scores + lrow * p.cand_stride;
            sink.gi = p.cand_ids + lrow * p.cand_stride;
            sink.cap = p.cand_stride;
            sink.appending = false;
            auto cursors_at = [&](int tile) {  // (re)position the CSR / exclusion cursors at the first object of `tile`
                const int64_t pos_first = (int64_t)tile * TILE_N + colg * COLS;
                const bool live = frow >= 0 && pos_first < p.n_pos;
                const int g_first = live ? (p.pos2obj ? __ldg(p.pos2obj + pos_first) : (int)pos_first) + p.id_off : 0;
                row_cursors_init(p, rs, live ? frow : -1, g_first);
                cw.cur = rs.cur;
                cw.fhi = rs.fhi;
                cw.streak = 0;
                window_load(p.indices, cw);
            };
            // (a macro, not a lambda: an outlined lambda would force every captured variable into local memory)
#define B200_STEP() fifo_step<QN, QS, WIDE>(p, rs, cw, qaddr, head, tail, ls, li, kc, sink)
            cursors_at(ts);
            int t = ts;
            uint32_t pos_t = (uint32_t)ts * TILE_N + (uint32_t)(colg * COLS);
            for (int it = 0; it < nt; ++it) {
                // exchange thresholds with the threads that own the other column groups of this row and with the other
                // ranks (monotone, racy by design: a stale value is only a weaker bound; the tag keeps a value of the
                // previous work item out)
                {
                    sts_thr(my_thr, work_tag, rs.thr);
                    if constexpr (NLIST == 2) {
                        uint32_t ptag;
                        float pthr;
                        lds_thr(other_thr, ptag, pthr);
                        if (ptag == work_tag) rs.thr = fmaxf(rs.thr, pthr);
                    } else {
#pragma unroll
                        for (int l = 0; l < NLIST; ++l) {  // (reading the own slot back is cheaper than a branch)
                            uint32_t ptag;
                            float pthr;
                            lds_thr(thr_row + (uint32_t)l * (TILE_M * 8), ptag, pthr);
                            if (ptag == work_tag) rs.thr = fmaxf(rs.thr, pthr);
                        }
                    }
                    if constexpr (PEERS) {
                        if (peers) {
                            uint32_t ptag;
                            float pthr;
                            lds_thr(thr_row + (uint32_t)NLIST * (TILE_M * 8), ptag, pthr);
                            if (ptag == work_tag) rs.thr = fmaxf(rs.thr, pthr);
                        }
                    }
                }
                if (WIDE && it == p.phase1_tiles) {
                    // wide mode: freeze the threshold.  Everything pending goes through the adaptive list first; then the
                    // list moves to the front of the row's global list and later candidates are appended behind it.
                    while (__any_sync(B200_FULL_MASK, head != tail)) B200_STEP();
                    if (row_ok) {
                        const int n = min(rs.cnt, kc);
                        for (int e = 0; e < n; ++e) {
                            sink.gs[e] = lds_f32(ls + e * 128);
                            sink.gi[e] = lds_s32(li + e * 128);
                        }
                    }
                    sink.appending = true;
                }
                mbar_wait(buf ? tfull1 : tfull0, tph);
                tc_fence_after();
                // the last tile before the stream wraps around / of the work item: every pending hit must be handled
                // before the cursors are repositioned or the list is written
                const bool force = (t + 1 == t1);
                const bool last = (it + 1 == nt);
                if (!dbg_skip) {
                    uint32_t r[COLS];
                    tmem_ld_sync(tbase + buf * (uint32_t)TILE_N, r);
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive_cluster(buf ? tempty1 : tempty0);  // accumulator free again
                    // (Measured and not kept: seeding the first tile's threshold with the K'-th largest group maximum -- the lists
                    // are empty there and every column is a hit -- changed nothing at N = 1M and gained 1.3 % on a 125 K-object
                    // shard, profiles/r02_ab_fused.txt.)
                    if constexpr (COLS == 128) {
                        const float m0 = chunk_max<0>(r), m1 = chunk_max<32>(r), m2 = chunk_max<64>(r), m3 = chunk_max<96>(r);
                        const float mx = fmaxf(max3(m0, m1, m2), m3);
                        if (__any_sync(B200_FULL_MASK, mx > rs.thr)) {
                            const float thr = rs.thr;
                            unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0;
                            if (__any_sync(B200_FULL_MASK, m0 > thr)) h0 = chunk_hits<0>(r, thr);
                            if (__any_sync(B200_FULL_MASK, m1 > thr)) h1 = chunk_hits<32>(r, thr);
                            if (__any_sync(B200_FULL_MASK, m2 > thr)) h2 = chunk_hits<64>(r, thr);
                            if (__any_sync(B200_FULL_MASK, m3 > thr)) h3 = chunk_hits<96>(r, thr);
                            for (;;) {
                                bool stuck = chunk_push<0, QN, QS>(r, h0, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) stuck = chunk_push<32, QN, QS>(r, h1, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) stuck = chunk_push<64, QN, QS>(r, h2, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) stuck = chunk_push<96, QN, QS>(r, h3, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) break;
                                B200_STEP();  // dense phase: make room, then go on
                            }
                        }
                    } else {
                        const float m0 = chunk_max<0>(r), m1 = chunk_max<32>(r);
                        const float mx = fmaxf(m0, m1);
                        if (__any_sync(B200_FULL_MASK, mx > rs.thr)) {
                            const float thr = rs.thr;
                            unsigned h0 = 0, h1 = 0;
                            if (__any_sync(B200_FULL_MASK, m0 > thr)) h0 = chunk_hits<0>(r, thr);
                            if (__any_sync(B200_FULL_MASK, m1 > thr)) h1 = chunk_hits<32>(r, thr);
                            for (;;) {
                                bool stuck = chunk_push<0, QN, QS>(r, h0, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) stuck = chunk_push<32, QN, QS>(r, h1, pos_t, rs.thr, n_pos, qaddr, head, tail);
                                if (!stuck) break;
                                B200_STEP();
                            }
                        }
                    }
                    // deferred work: at most one step per tile (bounded latency in front of the next accumulator), by default
                    // every PERIOD-th tile or as soon as some row has BACKLOG hits waiting (batching the steps measured +7 %),
                    // except where everything pending has to be finished
                    const bool due = (tail - head >= Cfg::BACKLOG) ||
                                     (head != tail && ((it & (Cfg::PERIOD - 1)) == Cfg::PERIOD - 1 || force || last));
                    if (__any_sync(B200_FULL_MASK, due)) {
                        B200_STEP();
                        if (force || last)
                            while (__any_sync(B200_FULL_MASK, head != tail)) B200_STEP();
                    }
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane0) mbar_arrive_cluster(buf ? tempty1 : tempty0);
                }
                buf ^= 1;
                tph ^= (buf == 0) ? 1u : 0u;
                ++t;
                pos_t += TILE_N;
                if (t == t1 && !last) {  // wrapped around: objects ascend again from the split's first tile
                    t = t0;
                    pos_t = (uint32_t)t0 * TILE_N + (uint32_t)(colg * COLS);
                    cursors_at(t0);
                }
            }
            // ---- this thread's candidate list (unsorted), its length and its final threshold
            if (row_ok) {
                if (!(WIDE && sink.appending)) {
                    const int n = min(rs.cnt, kc);
                    for (int e = 0; e < n; ++e) {
                        sink.gs[e] = lds_f32(ls + e * 128);
                        sink.gi[e] = lds_s32(li + e * 128);
                    }
                }
                p.cand_counts[lrow] = rs.cnt;
                p.cand_thr[lrow] = rs.thr;
            }
            // last work item of this pair: lets the helper warps leave their polling loop (kept INSIDE the loop: any code behind
            // it made ptxas spill the staged accumulator, 1.5 KB of stack)
            if (PEERS && peers && w + n_pairs >= n_work) sts_thr(my_thr, TAG_DONE, INFINITY);
        }
#undef B200_STEP
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
