#!/usr/bin/env python
"""bench.py -- recommend() users/sec of the B200 score + top-K engine on BASELINE.json's headline shape.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

A "step" is one pass of the hot path (score every user against the catalogue, mask viewed items, keep the K best)
over one batch of synthetic users.  Default workload = BASELINE config 2: |users| = |items| = 1M, d = 128, DOT, K = 10,
filter_viewed with 100 viewed items per user (SURVEY.md section 8d synthetic inputs: N(0,1)/sqrt(d) factors, fixed seeds).

  value  : whole-job users/sec with every input already resident in HBM (device-timed, max over ranks)
  e2e    : the same metric through the public host API (`Engine.topk` behind `B200Ranker`): per step the users'
           factors + CSR filter are copied from pinned host memory and the K (id, score) pairs copied back
  N > 1  : the catalogue is item-sharded over the ranks (north_star), every rank scores all users against its shard,
           one NCCL all-gather of U*K pairs + a merge kernel; total work is fixed => "scaling": "strong"
  --impl reference : the reference's CPU path (restatement of implicit.cpu.topk: BLAS sgemm + OpenMP select, all host
           threads) on a bounded sample of the same workload, rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 65536


def gen_factors(n, d, seed, lo=0, hi=None):
    """Rows [lo, hi) of the seeded N(0,1)/sqrt(d) fp32 matrix; generated in blocks so shards can be made independently."""
    hi = n if hi is None else hi
    out = np.empty((hi - lo, d), dtype=np.float32)
    b0 = lo // BLOCK
    pos = 0
    for b in range(b0, (hi + BLOCK - 1) // BLOCK):
        r0, r1 = b * BLOCK, min((b + 1) * BLOCK, n)
        blk = np.random.default_rng([seed, b]).standard_normal((r1 - r0, d), dtype=np.float32)
        blk *= np.float32(1.0 / np.sqrt(d))
        a, z = max(lo, r0), min(hi, r1)
        out[pos : pos + (z - a)] = blk[a - r0 : z - r0]
        pos += z - a
    return out


def gen_viewed(n_users, n_items, per_user, seed=2):
    """CSR of ~per_user viewed items per user: int64 indptr, int32 sorted indices (rare duplicates kept)."""
    cols = np.empty((n_users, per_user), dtype=np.int32)
    for b in range((n_users + BLOCK - 1) // BLOCK):
        r0, r1 = b * BLOCK, min((b + 1) * BLOCK, n_users)
        c = np.random.default_rng([seed, b]).integers(0, n_items, size=(r1 - r0, per_user), dtype=np.int32)
        c.sort(axis=1)
        cols[r0:r1] = c
    indptr = np.arange(n_users + 1, dtype=np.int64) * per_user
    return indptr, cols.reshape(-1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = (
        "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device):
        self.device, self.rows, self.proc, self.first = device, [], None, 0

    def mark(self):
        """Samples taken before this call (warm-up) are not reported.  The sampler is started BEFORE the warm-up because
        nvidia-smi's start-up (NVML initialisation over every GPU of the box) can stall CUDA calls of this process for tens
        of milliseconds -- measured as a one-off gap inside the first timed step when it was started right before it."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows[self.first:]:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except ValueError:
                continue
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "power_w_max": max(pw) if pw else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


def workload_name(a):
    return (
        f"config2: ImplicitALS-shaped factors, users={a.users} items={a.items} d={a.dim} Distance.{a.distance.upper()} "
        f"K={a.k} filter_viewed=True (~{a.viewed} viewed/user)"
    )


# --------------------------------------------------------------------------------------------------------------
def run_reference(a):
    """CPU arm: restatement of the reference's implicit.cpu.topk path on the host cores (rank 0 only)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import cpu_baseline
    from oracle.topk_oracle import calc_norms

    items = gen_factors(a.items, a.dim, 1)
    n_s = a.ref_users
    users = gen_factors(a.users, a.dim, 0, 0, min(a.users, n_s * (a.steps + a.warmup)))
    indptr, indices = gen_viewed(len(users), a.items, a.viewed)
    from scipy import sparse

    csr = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(len(users), a.items))
    norms = calc_norms(items) if a.distance == "cosine" else None
    threads = cpu_baseline.num_threads()
    times = []
    for s in range(a.warmup + a.steps):
        lo = (s * n_s) % max(1, len(users) - n_s + 1)
        t0 = time.perf_counter()
        cpu_baseline.topk_cpu(items, users[lo : lo + n_s], a.k, norms, csr[lo : lo + n_s], num_threads=0)
        dt = time.perf_counter() - t0
        if s >= a.warmup:
            times.append(dt)
    total = sum(times)
    value = n_s * len(times) / total
    line = {
        "impl": "reference",
        "metric": "recommend() users/sec",
        "value": value,
        "unit": "users/s",
        "n_gpus": a.gpus,
        "steps": a.steps,
        "warmup": a.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": f"{n_s} users per step against all {a.items} items"},
        "cpu_baseline": {
            "value": value, "unit": "users/s", "cores": threads, "kind": "port",
            "sample": f"{n_s} users x {a.items} items per step; numpy/OpenBLAS sgemm + C/OpenMP per-row select "
                      "(oracle/cpu_baseline.py, restating implicit.cpu.topk.topk as called at rank_implicit.py:264-272)",
        },
        "e2e": {"value": value, "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--viewed", type=int, default=100)
    ap.add_argument("--distance", default="dot", choices=["dot", "cosine"])
    ap.add_argument("--tc", default="auto", choices=["auto", "fp16", "bf16", "off"])
    ap.add_argument("--ref-users", type=int, default=1024, help="users per step of the CPU arms (bounded sample)")
    ap.add_argument("--parity-users", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--item-shards", type=int, default=0,
                    help="N > 1: item shards I (a divisor of N); the ranks form I item shards x N/I user groups.  0 = N (the north-star "
                         "scheme: every rank ranks all users against 1/N of the catalogue); 1 = plain user sharding")
    a = ap.parse_args()
    a.warmup = max(a.warmup, 0)

    if a.impl == "reference":
        run_reference(a)
        return

    import torch

    from rectools_b200 import Engine, _lib
    from rectools_b200.sharded import shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    # ---------------- partitioning: I item shards x Ug user groups (default I = world: item sharding, north star)
    n_ishards = a.item_shards if a.item_shards > 0 else world
    if world % n_ishards:
        raise SystemExit("--item-shards must divide the number of GPUs")
    n_ugroups = world // n_ishards
    shard_idx, group_idx = rank % n_ishards, rank // n_ishards
    ex_group = co_group = None  # exchange: my user group's item shards; collect: the holders of my item range
    if world > 1 and n_ugroups > 1:
        for g in range(n_ugroups):
            grp = dist.new_group([g * n_ishards + s_ for s_ in range(n_ishards)])
            if g == group_idx:
                ex_group = grp
        for s_ in range(n_ishards):
            grp = dist.new_group([g * n_ishards + s_ for g in range(n_ugroups)])
            if s_ == shard_idx:
                co_group = grp
    n_users_all = a.users
    wl_name = workload_name(a)
    per_group = -(-n_users_all // n_ugroups)
    u0, u1 = shard_bounds(n_users_all, n_ugroups)[group_idx]

    # ---------------- synthetic inputs (this rank's item range and user slice)
    lo, hi = shard_bounds(a.items, n_ishards)[shard_idx]
    items_local = gen_factors(a.items, a.dim, 1, lo, hi)
    users = gen_factors(n_users_all, a.dim, 0, u0, u1)
    indptr, indices = gen_viewed(n_users_all, a.items, a.viewed)
    if n_ugroups > 1:
        indices = indices[indptr[u0] : indptr[u1]].copy()
        indptr = (indptr[u0 : u1 + 1] - indptr[u0]).copy()
    a.users = u1 - u0  # rows this rank ranks; n_users_all is the whole job

    eng = Engine(items_local, cosine=a.distance == "cosine", device=local_rank, tc_mode=a.tc, id_offset=lo)
    info = eng.info()
    k = min(a.k, a.items)
    k_loc = min(k, hi - lo)

    # device-resident copies for the `value` measurement
    d_users = torch.from_numpy(users).to(dev)
    d_indptr = torch.from_numpy(indptr).to(dev)
    d_indices = torch.from_numpy(indices).to(dev)
    o_ids = torch.empty((a.users, k_loc), dtype=torch.int32, device=dev)
    o_sc = torch.empty((a.users, k_loc), dtype=torch.float32, device=dev)
    o_cnt = torch.empty((a.users,), dtype=torch.int32, device=dev)
    if world > 1:
        g_ids = torch.empty((n_ishards, a.users, k), dtype=torch.int32, device=dev)
        g_sc = torch.empty((n_ishards, a.users, k), dtype=torch.float32, device=dev)
        g_cnt = torch.empty((n_ishards, a.users), dtype=torch.int32, device=dev)
        # merged rows of my user group (allocated `per_group` long: the collect all-gather needs equal slices)
        m_ids = torch.full((per_group, k), -1, dtype=torch.int32, device=dev)
        m_sc = torch.zeros((per_group, k), dtype=torch.float32, device=dev)
        m_cnt = torch.zeros((per_group,), dtype=torch.int32, device=dev)
        if n_ugroups > 1:
            a_ids = torch.empty((n_ugroups * per_group, k), dtype=torch.int32, device=dev)
            a_sc = torch.empty((n_ugroups * per_group, k), dtype=torch.float32, device=dev)
            a_cnt = torch.empty((n_ugroups * per_group,), dtype=torch.int32, device=dev)
        else:
            a_ids, a_sc, a_cnt = m_ids, m_sc, m_cnt  # every rank already holds all rows
    lib = _lib.load()
    launches = [0]
    stats_log = []

    def exchange_and_merge():
        assert k_loc == k, "bench shards must hold at least k items"
        if n_ishards > 1:
            dist.all_gather_into_tensor(g_ids.view(n_ishards * a.users, k), o_ids, group=ex_group)
            dist.all_gather_into_tensor(g_sc.view(n_ishards * a.users, k), o_sc, group=ex_group)
            dist.all_gather_into_tensor(g_cnt.view(n_ishards * a.users), o_cnt, group=ex_group)
            _lib.check(lib.b200_rank_merge(local_rank, torch.cuda.current_stream().cuda_stream, n_ishards, a.users, k, g_ids.data_ptr(),
                                           g_sc.data_ptr(), g_cnt.data_ptr(), m_ids.data_ptr(), m_sc.data_ptr(), m_cnt.data_ptr()))
            launches[0] += 1 + (k + 31) // 32
        else:
            m_ids[: a.users].copy_(o_ids)
            m_sc[: a.users].copy_(o_sc)
            m_cnt[: a.users].copy_(o_cnt)
        if n_ugroups > 1:  # every rank receives the rows of the other user groups
            dist.all_gather_into_tensor(a_ids, m_ids, group=co_group)
            dist.all_gather_into_tensor(a_sc, m_sc, group=co_group)
            dist.all_gather_into_tensor(a_cnt, m_cnt, group=co_group)

    def step_resident():
        st = eng.topk_ptrs(
            a.users, k, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(),
            _lib.Q_INPUTS_ON_DEVICE | _lib.Q_OUTPUTS_ON_DEVICE,
            subjects=d_users.data_ptr(), indptr=d_indptr.data_ptr(), indices=d_indices.data_ptr(),
            stream=torch.cuda.current_stream().cuda_stream,
        )
        launches[0] += st["n_launches"]
        stats_log.append(st)
        if world > 1:
            exchange_and_merge()

    # pinned host buffers for the end-to-end measurement
    if not a.no_e2e:
        h_users = torch.from_numpy(users).pin_memory()
        h_indptr = torch.from_numpy(indptr).pin_memory()
        h_indices = torch.from_numpy(indices).pin_memory()
        h_ids = torch.empty((a.users, k_loc), dtype=torch.int32).pin_memory()
        h_sc = torch.empty((a.users, k_loc), dtype=torch.float32).pin_memory()
        h_cnt = torch.empty((a.users,), dtype=torch.int32).pin_memory()
        if world > 1:
            hm_ids = torch.empty(tuple(a_ids.shape), dtype=torch.int32).pin_memory()
            hm_sc = torch.empty(tuple(a_sc.shape), dtype=torch.float32).pin_memory()
    e2e_bytes = [0, 0]
    e2e_stats = {}

    def step_e2e():
        if world == 1:
            eng.topk(k, subjects=h_users.numpy(), indptr=h_indptr.numpy(), indices=h_indices.numpy(),
                     out=(h_ids.numpy(), h_sc.numpy(), h_cnt.numpy()))
            st = eng.last_stats
        else:
            st = eng.topk_ptrs(
                a.users, k, o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(), _lib.Q_OUTPUTS_ON_DEVICE,
                subjects=h_users.data_ptr(), indptr=h_indptr.data_ptr(), indices=h_indices.data_ptr(),
                stream=torch.cuda.current_stream().cuda_stream,
            )
            exchange_and_merge()
            if rank == 0:
                hm_ids.copy_(a_ids, non_blocking=True)
                hm_sc.copy_(a_sc, non_blocking=True)
                st = dict(st, d2h_bytes=int(a_ids.shape[0]) * k * 8)
            torch.cuda.current_stream().synchronize()
        e2e_bytes[0], e2e_bytes[1] = st["h2d_bytes"], st["d2h_bytes"]
        e2e_stats.clear()
        e2e_stats.update({kk: st[kk] for kk in ("ms_total", "ms_main", "ms_h2d", "ms_d2h")})

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    step_ms_log = []

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record()
        for i in range(steps):
            fn()
            ev[i + 1].record()
        barrier()
        ms = torch.tensor([ev[0].elapsed_time(ev[steps])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        step_ms_log.append([round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(steps)])  # this rank's per-step times
        return float(ms.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(a.warmup, 3) if a.steps > 0 else 0):
        step_resident()
    launches[0] = 0
    stats_log.clear()
    sampler.mark()
    total_ms = timed(step_resident, 0, a.steps)
    clocks = sampler.stop() if rank == 0 else None
    timed_launches = launches[0]
    timed_stats = list(stats_log)
    value = n_users_all * a.steps / (total_ms / 1e3)  # whole job: all user groups

    e2e = None
    if not a.no_e2e:
        e2e_steps = max(1, min(a.steps, 3))
        e2e_ms = timed(step_e2e, 1, e2e_steps)
        e2e = {
            "value": n_users_all * e2e_steps / (e2e_ms / 1e3), "unit": "users/s", "steps": e2e_steps,
            "h2d_bytes_per_step": int(e2e_bytes[0]), "d2h_bytes_per_step": int(e2e_bytes[1]),
            "api": "rectools_b200.Engine.topk (C ABI b200_rank_topk) with pinned host buffers",
            "engine_ms_last_step": dict(e2e_stats),
        }

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel (tensor-core candidate pass), timed by CUDA events in the engine
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        pass
    ms_main = float(np.mean([s["ms_main"] for s in timed_stats])) if timed_stats else float("nan")
    path = timed_stats[0]["path"] if timed_stats else -1
    n_loc = hi - lo
    flops = 2.0 * a.users * n_loc * a.dim
    if path == 1:
        peak = peaks.get("bf16_tflops_sustained")
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured; kernel runs ~all of a long step)"
        if peak is None:
            peak, peak_src = 1400.0, "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md; MEASURED_PEAKS.json absent)"
        achieved = flops / (ms_main * 1e-3) / 1e12
        # dram__bytes_read + write of one launch of this kernel from the committed `ncu --set full` capture, when that
        # capture was taken on exactly this workload (profiles/r01_ncu_tc_kernel.{txt,json}); otherwise null
        traffic = None
        try:
            cap = json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_tc_kernel.json")))
            if world == 1 and (cap["users"], cap["items"], cap["dim"]) == (a.users, a.items, a.dim):
                traffic = cap["dram_bytes_per_launch"]
        except (OSError, ValueError, KeyError):
            pass
        roof = {
            "kernel": "tc3_topk_kernel (TMA -> tcgen05.mma.cta_group::2 256x256x16 -> TMEM -> fused streaming top-K' selection)", "bound": "tensor",
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram read+write)", "peak_source": peak_src,
            "algorithmic": f"2*U*N_g*d = 2*{a.users}*{n_loc}*{a.dim} FLOP per launch",
            "ms_per_launch": ms_main,
            "peak_burst": peaks.get("bf16_tflops"),
            "item_stream": {
                "note": "item-factor HBM stream: one pass over the fp16 shard per wave of subject tiles",
                "shard_bytes": int(n_loc * info["d_pad"] * 2),
                "hbm_gbs_peak": peaks.get("hbm_gbs"),
            },
        }
    else:
        achieved = flops / (ms_main * 1e-3) / 1e12
        roof = {"kernel": "exact_topk_kernel", "bound": "fp64", "achieved": achieved, "peak": None, "unit": "TFLOP/s",
                "frac": None, "traffic": None, "ms_per_launch": ms_main}

    # ---------------- parity sample against the fp64 oracle, same run
    parity = None
    if a.parity_users > 0:
        from oracle.topk_oracle import rank_oracle
        from scipy import sparse

        # N > 1: rank 0 checks the merged result of (a sample of) its own user slice against the WHOLE catalogue
        n_par = a.parity_users if world == 1 else min(a.parity_users, 64)
        sel = np.linspace(0, a.users - 1, n_par).astype(np.int64)
        csr = sparse.csr_matrix((np.ones(len(indices), np.float32), indices, indptr), shape=(a.users, a.items))[sel]
        items_all = items_local if world == 1 else gen_factors(a.items, a.dim, 1)
        _, oid, osc = rank_oracle(a.distance, users, items_all, sel, k, csr, accum="f64")
        res_ids, res_sc = (o_ids, o_sc) if world == 1 else (a_ids, a_sc)  # (rank 0 = user group 0: its rows come first)
        got_ids = res_ids.cpu().numpy()[sel].reshape(-1)
        got_sc = res_sc.cpu().numpy()[sel].reshape(-1)
        if a.distance == "cosine":
            un = np.sqrt(np.einsum("ij,ij->i", users[sel], users[sel], dtype=np.float64)).astype(np.float32)
            osc = osc * np.repeat(un, k)
        parity = {
            "users_checked": int(len(sel)), "id_mismatches": int((got_ids != oid).sum()),
            "max_rel_score_err": float(np.max(np.abs(got_sc - osc) / np.maximum(np.abs(osc), 1e-30))),
            "oracle": "oracle/topk_oracle.py rank_oracle(accum='f64')",
        }

    cpu = None
    if not a.no_cpu_baseline and world == 1:
        from oracle import cpu_baseline
        from scipy import sparse

        n_s = min(a.ref_users, a.users)
        csr = sparse.csr_matrix((np.ones(n_s * a.viewed, np.float32), indices[: n_s * a.viewed], indptr[: n_s + 1]),
                                shape=(n_s, a.items))
        cpu_baseline.topk_cpu(items_local, users[: min(64, n_s)], k, None, csr[: min(64, n_s)])  # warm-up
        t0 = time.perf_counter()
        reps = 0
        while reps < 1 or (time.perf_counter() - t0 < 10 and reps < 8):
            cpu_baseline.topk_cpu(items_local, users[:n_s], k, None, csr)
            reps += 1
        dt = (time.perf_counter() - t0) / reps
        cpu = {"value": n_s / dt, "unit": "users/s", "cores": cpu_baseline.num_threads(), "kind": "port",
               "sample": f"{n_s} users x {a.items} items, {reps} repetitions; numpy/OpenBLAS sgemm + C/OpenMP select "
                         "restating implicit.cpu.topk.topk (rank_implicit.py:264-272)"}

    line = {
        "metric": "recommend() users/sec",
        "value": value,
        "unit": "users/s",
        "n_gpus": world,
        "steps": a.steps,
        "warmup": max(a.warmup, 3),
        "ms_per_step": total_ms / max(a.steps, 1),
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": ("f16" if info["tc_dtype"] == 1 else "bf16" if info["tc_dtype"] == 2 else "f64")
        + " tensor-core candidates + f64-accumulated f32 re-score",
        "data": "synthetic",
        "config": {
            "workload": wl_name,
            "parallelism": (
                "single GPU" if world == 1 else
                f"items sharded over {world} GPU(s), NCCL all-gather + merge" if n_ugroups == 1 else
                f"users sharded over {world} GPU(s), NCCL all-gather of the results" if n_ishards == 1 else
                f"grid: {n_ishards} item shards x {n_ugroups} user groups, NCCL all-gather + merge per user group, all-gather of the results"
            ),
            "l2": "inputs larger than L2 (fp16 item shard %.0f MB + users %.0f MB per step)"
            % (n_loc * info["d_pad"] * 2 / 1e6, a.users * info["d_pad"] * 2 / 1e6),
            "engine": {kk: timed_stats[0][kk] for kk in ("path", "k_cand", "n_splits", "n_fallback_rows", "n_exact_rows")} if timed_stats else {},
            "engine_ms_last_step": {kk: timed_stats[-1][kk] for kk in ("ms_total", "ms_main", "ms_h2d", "ms_d2h")} if timed_stats else {},
            "device": info["device_name"],
        },
        "e2e": e2e,
        "gpu_launches": int(timed_launches),
        "ms_steps_rank0": step_ms_log[0] if step_ms_log else None,
        "clocks": clocks,
        "roofline": roof,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
