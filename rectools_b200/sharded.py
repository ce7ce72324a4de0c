"""Sharded ranking over several GPUs (one process per GPU, `torch.distributed`).

The reference is single-device.  Default = the north-star multi-GPU scheme (section 8e): the object
catalogue is split into contiguous ranges, every rank scores ALL subjects against its range and keeps a local top-k
with GLOBAL object ids, the ranks exchange `n_rows * k` (id, score) pairs with one all-gather (NCCL over NVLink on
GPUs) and every rank merges the `world * k` candidates per subject (`b200_rank_merge`).  Exact local lists => exact
global top-k; ties resolve by (score desc, id asc) in the merge exactly as inside a shard.

`item_shards=I` (a divisor of the world size) selects the other partitionings of section 8e: the ranks form a grid of
I item shards x world/I subject groups; a rank scores ITS slice of the subject batch against ITS item range, the I ranks
of a subject group exchange + merge as above, and one more all-gather among the ranks holding the same item range hands
every rank the rows of the other subject groups.  `item_shards=1` is plain subject sharding (no merge at all).  Larger
item ranges keep the fused kernel in its efficient regime (617 TFLOP/s at 125 K items per GPU,
1030 at 1 M), smaller ones are what a catalogue that does not fit one GPU needs.
"""
from __future__ import annotations

import typing as tp

import numpy as np
from scipy import sparse

from .ranker import Distance, Engine, _as_distance, _dense_f32, flatten_padded, prepare_factors


def shard_bounds(n_objects: int, world_size: int) -> tp.List[tp.Tuple[int, int]]:
    """Contiguous ranges of ceil(n/world) objects (the last ones may be short or empty)."""
    per = -(-n_objects // world_size) if world_size > 0 else n_objects
    return [(min(r * per, n_objects), min((r + 1) * per, n_objects)) for r in range(world_size)]


def split_whitelist(whitelist: np.ndarray, lo: int, hi: int) -> np.ndarray:
    """Part of a sorted global whitelist that falls into [lo, hi), as LOCAL positions of that shard."""
    whitelist = np.asarray(whitelist, dtype=np.int64)
    a, b = np.searchsorted(whitelist, [lo, hi], side="left")
    return (whitelist[a:b] - lo).astype(np.int32)


def merge_padded_numpy(ids: np.ndarray, scores: np.ndarray, counts: np.ndarray, k: int):
    """Host restatement of `b200_rank_merge` (used by the CPU/gloo tests of the exchange logic only)."""
    n_lists, n_rows, _ = ids.shape
    out_ids = np.full((n_rows, k), -1, dtype=np.int32)
    out_sc = np.full((n_rows, k), -np.finfo(np.float32).max, dtype=np.float32)
    out_cnt = np.zeros(n_rows, dtype=np.int32)
    for r in range(n_rows):
        ci = np.concatenate([ids[l, r, : counts[l, r]] for l in range(n_lists)])
        cs = np.concatenate([scores[l, r, : counts[l, r]] for l in range(n_lists)])
        order = np.lexsort((ci, -cs.astype(np.float64)))[:k]
        out_ids[r, : len(order)] = ci[order]
        out_sc[r, : len(order)] = cs[order]
        out_cnt[r] = len(order)
    return out_ids, out_sc, out_cnt


class EngineShard:
    """Local top-k provider backed by the CUDA engine (device tensors in / out)."""

    def __init__(self, objects: np.ndarray, cosine: bool, lo: int, device: int, tc_mode: str) -> None:
        import torch

        self.torch = torch
        self.device = torch.device("cuda", device)
        self.engine = Engine(objects, cosine=cosine, device=device, tc_mode=tc_mode, id_offset=lo)

    def set_subjects(self, subjects: np.ndarray) -> None:
        self.engine.set_subjects(subjects)

    def local_topk(self, subject_ids, k, indptr, indices, whitelist_local):
        from . import _lib

        torch = self.torch
        n = len(subject_ids)
        n_pos_local = self.engine.n_objects if whitelist_local is None else len(whitelist_local)
        k_loc = min(k, n_pos_local)  # the engine writes rows of k_out = min(k, local candidates) columns
        ids = torch.empty((n, k_loc), dtype=torch.int32, device=self.device)
        sc = torch.empty((n, k_loc), dtype=torch.float32, device=self.device)
        cnt = torch.zeros((n,), dtype=torch.int32, device=self.device)
        if k_loc == 0 or n == 0:
            return ids, sc, cnt
        keep = [np.ascontiguousarray(subject_ids, dtype=np.int64)]
        kw = dict(subject_ids=keep[0].ctypes.data)
        if indptr is not None:
            keep += [np.ascontiguousarray(indptr, dtype=np.int64), np.ascontiguousarray(indices, dtype=np.int32)]
            kw.update(indptr=keep[1].ctypes.data, indices=keep[2].ctypes.data)
        if whitelist_local is not None:
            keep.append(np.ascontiguousarray(whitelist_local, dtype=np.int32))
            kw.update(whitelist=keep[-1].ctypes.data, n_whitelist=len(keep[-1]))
        self.engine.topk_ptrs(
            n, k, ids.data_ptr(), sc.data_ptr(), cnt.data_ptr(), _lib.Q_OUTPUTS_ON_DEVICE,
            stream=torch.cuda.current_stream().cuda_stream, **kw,
        )
        return ids, sc, cnt

    def merge(self, ids, sc, cnt, k):
        from . import _lib

        torch = self.torch
        n_lists, n = ids.shape[0], ids.shape[1]
        o_ids = torch.empty((n, k), dtype=torch.int32, device=self.device)
        o_sc = torch.empty((n, k), dtype=torch.float32, device=self.device)
        o_cnt = torch.empty((n,), dtype=torch.int32, device=self.device)
        _lib.check(
            _lib.load().b200_rank_merge(
                self.device.index, torch.cuda.current_stream().cuda_stream, n_lists, n, k, ids.data_ptr(), sc.data_ptr(),
                cnt.data_ptr(), o_ids.data_ptr(), o_sc.data_ptr(), o_cnt.data_ptr(),
            )
        )
        return o_ids, o_sc, o_cnt


class ShardedB200Ranker:
    """`Ranker`-protocol object whose catalogue is sharded over the ranks of a `torch.distributed` group.

    `objects_factors` is the FULL matrix (every rank slices its own range) unless `objects_are_local=True`, in which case
    it is this rank's range and `n_objects_total` must be given.  Every rank must call `rank()` with the same arguments
    (SPMD); every rank returns the full, identical result.
    `local_factory(objects_local, cosine, lo)` may replace the CUDA engine with another local top-k provider -- the
    CPU (gloo) tests of the exchange logic plug the oracle in here.
    """

    def __init__(
        self,
        distance: tp.Any,
        subjects_factors: tp.Any,
        objects_factors: tp.Any,
        group: tp.Any = None,
        device: tp.Optional[int] = None,
        tc_mode: str = "auto",
        objects_are_local: bool = False,
        n_objects_total: tp.Optional[int] = None,
        local_factory: tp.Optional[tp.Callable[..., tp.Any]] = None,
        item_shards: tp.Optional[int] = None,
    ) -> None:
        import torch
        import torch.distributed as dist

        self.dist, self.torch, self.group = dist, torch, group
        self.rank_id, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.item_shards = self.world if item_shards is None else int(item_shards)
        if self.item_shards < 1 or self.world % self.item_shards:
            raise ValueError("`item_shards` must divide the world size")
        self.subject_groups = self.world // self.item_shards
        self.shard_idx, self.group_idx = self.rank_id % self.item_shards, self.rank_id // self.item_shards
        # communicators: `exchange` = the ranks of my subject group (one per item shard), `collect` = the ranks that hold my
        # item range (one per subject group).  new_group is collective over the parent group: every rank creates all of them.
        self.exchange_group, self.collect_group = group, None
        if self.subject_groups > 1:
            ranks = list(range(self.world)) if group is None else dist.get_process_group_ranks(group)
            for g in range(self.subject_groups):
                grp = dist.new_group([ranks[g * self.item_shards + s] for s in range(self.item_shards)])
                if g == self.group_idx:
                    self.exchange_group = grp
            for s_ in range(self.item_shards):
                grp = dist.new_group([ranks[g * self.item_shards + s_] for g in range(self.subject_groups)])
                if s_ == self.shard_idx:
                    self.collect_group = grp
        self.distance = _as_distance(distance)
        subjects = _dense_f32(subjects_factors)
        objects = _dense_f32(objects_factors)
        n_total = int(n_objects_total) if objects_are_local else objects.shape[0]
        self.bounds = shard_bounds(n_total, self.item_shards)
        self.lo, self.hi = self.bounds[self.shard_idx]
        if not objects_are_local:
            objects = objects[self.lo : self.hi]
        if objects.shape[0] != self.hi - self.lo:
            raise ValueError("local object matrix does not match this rank's shard range")
        self.n_subjects, self.n_objects = subjects.shape[0], n_total
        subjects, objects, self.subjects_norms, self.subjects_dots = prepare_factors(self.distance, subjects, objects)
        cosine = self.distance == Distance.COSINE
        if local_factory is not None:
            self.local = local_factory(objects, cosine, self.lo)
        else:
            dev = torch.cuda.current_device() if device is None else device
            self.local = EngineShard(objects, cosine, self.lo, dev, tc_mode)
        self.local.set_subjects(subjects)

    def rank_padded(self, subject_ids, k=None, filter_pairs_csr=None, sorted_object_whitelist=None):
        subject_ids = np.asarray(subject_ids, dtype=np.int64).reshape(-1)
        if filter_pairs_csr is not None and filter_pairs_csr.shape[0] != len(subject_ids):
            raise ValueError("Number of rows in `filter_pairs_csr` must be equal to `len(sublect_ids)`")
        wl_local = None
        n_pos_total = self.n_objects
        n_pos_local = self.hi - self.lo
        if sorted_object_whitelist is not None:
            wl_local = split_whitelist(sorted_object_whitelist, self.lo, self.hi)
            n_pos_total, n_pos_local = len(sorted_object_whitelist), len(wl_local)
        if k is None:
            k = n_pos_total
        k = min(int(k), n_pos_total)
        if k <= 0:
            raise ValueError("`k` must be positive")
        indptr = indices = None
        if filter_pairs_csr is not None:
            csr = filter_pairs_csr if sparse.isspmatrix_csr(filter_pairs_csr) else sparse.csr_matrix(filter_pairs_csr)
            if not csr.has_sorted_indices:
                csr = csr.sorted_indices()
            indptr, indices = csr.indptr, csr.indices
        torch = self.torch
        n_all = len(subject_ids)
        # my subject group's slice of the batch (contiguous rows; the CSR filter is sliced by its row pointer)
        row_bounds = shard_bounds(n_all, self.subject_groups)
        r0, r1 = row_bounds[self.group_idx]
        my_ids = subject_ids[r0:r1]
        my_indptr = my_indices = None
        if indptr is not None:
            my_indptr = np.asarray(indptr[r0 : r1 + 1], dtype=np.int64) - int(indptr[r0])
            my_indices = indices[int(indptr[r0]) : int(indptr[r1])]
        ids, sc, cnt = self.local.local_topk(my_ids, k, my_indptr, my_indices, wl_local)
        n = len(my_ids)
        k_loc = ids.shape[1]
        if k_loc < k:  # short shard: pad to the common width
            pad_i = torch.full((n, k), -1, dtype=ids.dtype, device=ids.device)
            pad_s = torch.full((n, k), -3.4028234663852886e38, dtype=sc.dtype, device=sc.device)
            pad_i[:, :k_loc], pad_s[:, :k_loc] = ids, sc
            ids, sc = pad_i, pad_s
        if self.item_shards > 1:
            w = self.item_shards
            g_ids = torch.empty((w, n, k), dtype=ids.dtype, device=ids.device)
            g_sc = torch.empty((w, n, k), dtype=sc.dtype, device=sc.device)
            g_cnt = torch.empty((w, n), dtype=cnt.dtype, device=cnt.device)
            # (concatenated-along-dim-0 views: the form every backend accepts)
            self.dist.all_gather_into_tensor(g_ids.view(w * n, k), ids.contiguous(), group=self.exchange_group)
            self.dist.all_gather_into_tensor(g_sc.view(w * n, k), sc.contiguous(), group=self.exchange_group)
            self.dist.all_gather_into_tensor(g_cnt.view(w * n), cnt.contiguous(), group=self.exchange_group)
            o_ids, o_sc, o_cnt = self.local.merge(g_ids, g_sc, g_cnt, k)
        else:
            o_ids, o_sc, o_cnt = ids, sc, cnt
        if self.subject_groups > 1:
            # hand every rank the rows of the other subject groups (slices padded to the longest one)
            u = self.subject_groups
            per = max(b - a for a, b in row_bounds)
            def padded(t, fill):
                if t.shape[0] == per:
                    return t.contiguous()
                out = torch.full((per,) + tuple(t.shape[1:]), fill, dtype=t.dtype, device=t.device)
                out[: t.shape[0]] = t
                return out
            a_ids = torch.empty((u * per, k), dtype=o_ids.dtype, device=o_ids.device)
            a_sc = torch.empty((u * per, k), dtype=o_sc.dtype, device=o_sc.device)
            a_cnt = torch.empty((u * per,), dtype=o_cnt.dtype, device=o_cnt.device)
            self.dist.all_gather_into_tensor(a_ids, padded(o_ids, -1), group=self.collect_group)
            self.dist.all_gather_into_tensor(a_sc, padded(o_sc, -3.4028234663852886e38), group=self.collect_group)
            self.dist.all_gather_into_tensor(a_cnt, padded(o_cnt, 0), group=self.collect_group)
            keep = torch.cat([torch.arange(g * per, g * per + (b - a), device=a_ids.device) for g, (a, b) in enumerate(row_bounds)])
            o_ids, o_sc, o_cnt = a_ids[keep], a_sc[keep], a_cnt[keep]
        return subject_ids, o_ids, o_sc, o_cnt

    def rank(self, subject_ids, k=None, filter_pairs_csr=None, sorted_object_whitelist=None):
        subject_ids, ids, sc, cnt = self.rank_padded(subject_ids, k, filter_pairs_csr, sorted_object_whitelist)
        ids, sc, cnt = (t.cpu().numpy() if hasattr(t, "cpu") else np.asarray(t) for t in (ids, sc, cnt))
        all_subjects, all_ids, all_scores = flatten_padded(subject_ids, ids, sc, cnt)
        if self.distance == Distance.COSINE:
            all_scores = all_scores / self.subjects_norms[all_subjects]
        elif self.distance == Distance.EUCLIDEAN:
            all_scores = np.sqrt(np.maximum(self.subjects_dots[all_subjects] - all_scores, 0)).astype(np.float32)
        return all_subjects, all_ids, all_scores
