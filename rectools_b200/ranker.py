"""`B200Ranker`: the reference's `Ranker` protocol on top of libb200rank.so.

Mirrors `rectools.models.rank.ImplicitRanker` (rectools/models/rank/rank_implicit.py:34-280) -- same constructor
meaning `(distance, subjects_factors, objects_factors, ...)`, same `rank(subject_ids, k, filter_pairs_csr,
sorted_object_whitelist)` signature, return triplet and error behaviour -- with the native top-k call
(rank_implicit.py:264-272 / :175-182) and the per-user Python post-loop (:120-146) replaced by one C-ABI call plus
vectorised numpy.  There is no CPU fallback: construction fails if the CUDA library or an sm_100 device is missing.
"""
from __future__ import annotations

import ctypes as C
import typing as tp
import weakref
from enum import Enum

import numpy as np
from scipy import sparse

from . import _lib

InternalIds = tp.Sequence[int]
Scores = tp.Union[tp.Sequence[float], np.ndarray]


class Distance(str, Enum):
    """Same members / values as `rectools.models.rank.Distance` (rectools/models/rank/rank.py:25-30)."""

    DOT = "dot"
    COSINE = "cosine"
    EUCLIDEAN = "euclidean"


_TC_MODES = {"auto": _lib.TC_AUTO, "fp16": _lib.TC_FP16, "bf16": _lib.TC_BF16, "off": _lib.TC_OFF}


def _as_distance(distance: tp.Any) -> Distance:
    return Distance(str(getattr(distance, "value", distance)))


def _dense_f32(x: tp.Any) -> np.ndarray:
    """`factors.astype(np.float32)` of the reference (rank_implicit.py:70-71), C-contiguous, torch tensors accepted."""
    if sparse.issparse(x):
        raise TypeError("sparse factors are kept sparse (B200Ranker: CSR subjects), never densified as a whole")
    if hasattr(x, "detach") and hasattr(x, "cpu"):
        x = x.detach().cpu().numpy()
    x = np.asarray(x)
    if x.ndim != 2:
        raise ValueError("factor matrices must be 2-dimensional")
    return np.ascontiguousarray(x, dtype=np.float32)


def _is_cuda_tensor(x: tp.Any) -> bool:
    return hasattr(x, "is_cuda") and bool(getattr(x, "is_cuda")) and hasattr(x, "data_ptr")


def _norms_f32(x: np.ndarray) -> np.ndarray:
    """`_calc_norms(avoid_zeros=True)` (rank_implicit.py:98-105), accumulated in fp64 as the engine does."""
    n = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float64)).astype(np.float32)
    n[n == 0] = 1e-10
    return n


def check_whitelist(whitelist: np.ndarray, n_objects: int) -> None:
    """`sorted_object_whitelist` (rank.py:39): object ids in range, strictly ascending -- the kernels merge a row's viewed ids
    against the whitelist positions in ascending order, so an unsorted whitelist would let viewed objects through."""
    if len(whitelist) == 0:
        return
    if whitelist[0] < 0 or whitelist[-1] >= n_objects or whitelist.min() < 0 or whitelist.max() >= n_objects:
        raise IndexError("whitelist id out of range")
    if len(whitelist) > 1 and not bool((np.diff(whitelist) > 0).all()):
        raise ValueError("`sorted_object_whitelist` must be sorted ascending without duplicates")


def prepare_factors(
    distance: Distance, subjects: np.ndarray, objects: np.ndarray
) -> tp.Tuple[np.ndarray, np.ndarray, tp.Optional[np.ndarray], tp.Optional[np.ndarray]]:
    """Host prologue of `ImplicitRanker`: COSINE subject norms (rank_implicit.py:76-77) or the EUCLIDEAN -> DOT
    augmentation (rank_implicit.py:79-81, :242-246).  Returns (subjects', objects', subjects_norms, subjects_dots)."""
    norms = dots = None
    if distance == Distance.COSINE:
        norms = _norms_f32(subjects)
    elif distance == Distance.EUCLIDEAN:
        dots = (subjects**2).sum(axis=1)
        subjects = np.hstack((-np.ones((subjects.shape[0], 1)), 2 * subjects)).astype(np.float32)
        objects = np.hstack(((objects**2).sum(axis=1).reshape(-1, 1), objects)).astype(np.float32)
    return subjects, objects, norms, dots


class Engine:
    """Owner of one `b200_rank_engine*` (resident object factors on one GPU)."""

    def __init__(
        self,
        objects: np.ndarray,
        cosine: bool,
        device: int = 0,
        tc_mode: str = "auto",
        id_offset: int = 0,
        objects_device_ptr: tp.Optional[int] = None,
        shape: tp.Optional[tp.Tuple[int, int]] = None,
        objects_dtype: int = _lib.DT_F32,
    ) -> None:
        self._lib = _lib.load()
        self._h = C.c_void_p()
        if objects_device_ptr is not None:
            assert shape is not None
            n, d = shape
            ptr, flags = objects_device_ptr, _lib.F_OBJECTS_ON_DEVICE
            self._keep = None
        else:
            objects = np.ascontiguousarray(objects, dtype=np.float32)
            n, d = objects.shape
            ptr, flags = objects.ctypes.data, 0
            self._keep = objects
        _lib.check(
            self._lib.b200_rank_create_ex(
                C.byref(self._h), ptr, objects_dtype, n, d, _lib.DIST_COSINE if cosine else _lib.DIST_DOT, device,
                _TC_MODES[tc_mode], flags,
            )
        )
        self._keep = None  # the engine copied the host matrix
        self.n_objects, self.d, self.device = int(n), int(d), int(device)
        if id_offset:
            _lib.check(self._lib.b200_rank_set_id_offset(self._h, int(id_offset)))
        self.id_offset = int(id_offset)
        self.last_stats: tp.Dict[str, tp.Any] = {}

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.b200_rank_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self) -> None:  # pragma: no cover
        try:
            self.close()
        except Exception:  # pylint: disable=broad-except
            pass

    def info(self) -> tp.Dict[str, tp.Any]:
        inf = _lib.Info()
        _lib.check(self._lib.b200_rank_get_info(self._h, C.byref(inf)))
        out = {name: getattr(inf, name) for name, _ in inf._fields_}  # pylint: disable=protected-access
        out["device_name"] = inf.device_name.decode()
        return out

    def set_subjects(self, subjects: np.ndarray, key: tp.Optional[tp.Hashable] = None, owner: tp.Any = None) -> None:
        """Upload the resident subject factors.  `key` (optional identity of the matrix): a repeated call with the key of the
        matrix that is already resident is a no-op -- `VectorModel` builds a new ranker per `recommend()` call (vector.py:66).
        `owner`: the ranker whose subjects are resident now (several rankers may share one cached engine)."""
        subjects = np.ascontiguousarray(subjects, dtype=np.float32)
        if subjects.shape[1] != self.d:
            raise ValueError("subject and object factors must have the same number of columns")
        self._subjects_owner = weakref.ref(owner) if owner is not None else None
        if key is not None and key == getattr(self, "_subjects_key", None):
            return
        _lib.check(self._lib.b200_rank_set_subjects(self._h, subjects.ctypes.data, subjects.shape[0], 0))
        self._subjects_key = key

    @property
    def subjects_owner(self) -> tp.Any:
        """The ranker whose subject factors are resident (None: nobody's / collected).  A weak reference: a cached engine
        must not keep its last ranker and that ranker's matrices alive."""
        ref = getattr(self, "_subjects_owner", None)
        return ref() if ref is not None else None

    def set_subjects_device(self, ptr: int, n_subjects: int) -> None:
        _lib.check(self._lib.b200_rank_set_subjects(self._h, ptr, n_subjects, 1))

    def peer_export(self, max_rows: int) -> bytes:
        """Allocate this engine's published-threshold array (threshold sharing between the ranks of an item-sharded
        catalogue) and return its 64-byte CUDA IPC handle."""
        buf = C.create_string_buffer(64)
        _lib.check(self._lib.b200_rank_peer_export(self._h, int(max_rows), buf))
        return buf.raw

    def peer_import(self, handles: tp.Sequence[bytes], self_index: int) -> None:
        """Open the published-threshold arrays of all ranks (`handles` in rank order, this engine's own included)."""
        blob = b"".join(handles)
        assert len(blob) == 64 * len(handles)
        _lib.check(self._lib.b200_rank_peer_import(self._h, len(handles), int(self_index), blob))

    def topk_raw(self, q: _lib.Query) -> tp.Dict[str, tp.Any]:
        st = _lib.Stats()
        _lib.check(self._lib.b200_rank_topk(self._h, C.byref(q), C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def topk_ptrs(
        self,
        n_rows: int,
        k: int,
        out_ids: int,
        out_scores: int,
        out_counts: int,
        flags: int,
        subjects: int = 0,
        subject_ids: int = 0,
        n_subjects_total: int = 0,
        indptr: int = 0,
        indices: int = 0,
        whitelist: int = 0,
        n_whitelist: int = 0,
        stream: int = 0,
        out_bounds: int = 0,
        peer_epoch: int = 0,
        subject_dtype: int = _lib.DT_F32,
    ) -> tp.Dict[str, tp.Any]:
        """Raw-pointer call (host or device addresses according to `flags`); returns the call statistics."""
        q = _lib.Query()
        q.out_bounds, q.peer_epoch, q.subject_dtype = out_bounds or None, int(peer_epoch), int(subject_dtype)
        q.subjects, q.subject_ids, q.n_rows, q.n_subjects_total = subjects or None, subject_ids or None, n_rows, n_subjects_total
        q.csr_indptr, q.csr_indices = indptr or None, indices or None
        q.whitelist, q.n_whitelist = whitelist or None, n_whitelist
        q.k, q.flags = int(k), int(flags)
        q.out_ids, q.out_scores, q.out_counts = out_ids, out_scores, out_counts
        q.stream = stream or None
        return self.topk_raw(q)

    def topk(
        self,
        k: int,
        subjects: tp.Optional[np.ndarray] = None,
        subject_ids: tp.Optional[np.ndarray] = None,
        indptr: tp.Optional[np.ndarray] = None,
        indices: tp.Optional[np.ndarray] = None,
        whitelist: tp.Optional[np.ndarray] = None,
        flags: int = 0,
        out: tp.Optional[tp.Tuple[np.ndarray, np.ndarray, np.ndarray]] = None,
        sparse_subjects: tp.Optional[sparse.csr_matrix] = None,
    ) -> tp.Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Host-buffer call: returns padded `(ids [n,k_out] int32, scores [n,k_out] fp32, counts [n] int32)`.
        `sparse_subjects`: the batch rows as a CSR matrix [n_rows, d] (EASE), instead of `subjects` / `subject_ids`."""
        q = _lib.Query()
        keep = []
        if sparse_subjects is not None:
            if subjects is not None or subject_ids is not None:
                raise ValueError("sparse_subjects excludes subjects / subject_ids")
            if sparse_subjects.shape[1] != self.d:
                raise ValueError("subject and object factors must have the same number of columns")
            sp_indptr = np.ascontiguousarray(sparse_subjects.indptr, dtype=np.int64)
            sp_indices = np.ascontiguousarray(sparse_subjects.indices, dtype=np.int32)
            sp_data = np.ascontiguousarray(sparse_subjects.data, dtype=np.float32)
            q.sub_indptr, q.sub_indices, q.sub_data = sp_indptr.ctypes.data, sp_indices.ctypes.data, sp_data.ctypes.data
            keep += [sp_indptr, sp_indices, sp_data]
        if subjects is not None:
            subjects = np.ascontiguousarray(subjects, dtype=np.float32)
            if subjects.ndim != 2 or subjects.shape[1] != self.d:
                raise ValueError("subject and object factors must have the same number of columns")
            q.subjects = subjects.ctypes.data
            keep.append(subjects)
        if subject_ids is not None:
            subject_ids = np.ascontiguousarray(subject_ids, dtype=np.int64)
            q.subject_ids = subject_ids.ctypes.data
            keep.append(subject_ids)
            n_rows = len(subject_ids)
            q.n_subjects_total = 0 if subjects is None else subjects.shape[0]
        elif sparse_subjects is not None:
            n_rows = sparse_subjects.shape[0]
        else:
            if subjects is None:
                raise ValueError("either subjects or subject_ids is required")
            n_rows = subjects.shape[0]
        q.n_rows = n_rows
        if indptr is not None:
            indptr = np.ascontiguousarray(indptr, dtype=np.int64)
            if len(indptr) != n_rows + 1:
                raise ValueError("Number of rows in `filter_pairs_csr` must be equal to `len(sublect_ids)`")
            indices = np.ascontiguousarray(indices if indices is not None else np.empty(0), dtype=np.int32)
            q.csr_indptr = indptr.ctypes.data
            q.csr_indices = indices.ctypes.data
            keep += [indptr, indices]
        n_pos = self.n_objects
        if whitelist is not None:
            whitelist = np.ascontiguousarray(whitelist, dtype=np.int32)
            q.whitelist = whitelist.ctypes.data
            q.n_whitelist = len(whitelist)
            n_pos = len(whitelist)
            keep.append(whitelist)
        q.k = int(k)
        q.flags = int(flags)
        k_out = max(0, min(int(k), n_pos))
        if out is None:
            ids = np.empty((n_rows, k_out), dtype=np.int32)
            scores = np.empty((n_rows, k_out), dtype=np.float32)
            counts = np.zeros(n_rows, dtype=np.int32)
        else:
            ids, scores, counts = out
        q.out_ids, q.out_scores, q.out_counts = ids.ctypes.data, scores.ctypes.data, counts.ctypes.data
        self.topk_raw(q)
        del keep
        return ids, scores, counts


def flatten_padded(
    subject_ids: np.ndarray, ids: np.ndarray, scores: np.ndarray, counts: np.ndarray
) -> tp.Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Vectorised `_process_implicit_scores` (rank_implicit.py:120-146): padded rows -> flat ragged triplet."""
    k_out = ids.shape[1] if ids.ndim == 2 else 0
    if k_out == 0 or len(subject_ids) == 0:
        return np.empty(0, dtype=np.int64), np.empty(0, dtype=np.int64), np.empty(0, dtype=np.float32)
    if int(counts.min()) == k_out:
        return np.repeat(subject_ids, k_out), ids.reshape(-1).astype(np.int64), scores.reshape(-1)
    mask = np.arange(k_out, dtype=np.int32)[None, :] < counts[:, None]
    return np.repeat(subject_ids, counts), ids[mask].astype(np.int64), scores[mask]


class B200Ranker:
    """Ranker backed by the B200 engine.

    Parameters mirror `ImplicitRanker.__init__` (rank_implicit.py:58-65); `num_threads` / `use_gpu` are accepted and
    ignored so that the class can be bound in place of `ImplicitRanker` (rectools/models/vector.py:66-72).

    Parameters
    ----------
    distance : Distance | str
    subjects_factors : np.ndarray | scipy.sparse.csr_matrix | torch.Tensor, shape (n_subjects, n_factors)
    objects_factors : np.ndarray | torch.Tensor, shape (n_objects, n_factors)
    device : int, CUDA device ordinal
    tc_mode : "auto" | "fp16" | "bf16" | "off" -- dtype of the tensor-core candidate pass ("off": fp64 kernel only)
    """

    def __init__(
        self,
        distance: tp.Any,
        subjects_factors: tp.Any,
        objects_factors: tp.Any,
        num_threads: int = 0,  # pylint: disable=unused-argument
        use_gpu: bool = True,  # pylint: disable=unused-argument
        device: int = 0,
        tc_mode: str = "auto",
        engine: tp.Optional[Engine] = None,
        subjects_key: tp.Optional[tp.Hashable] = None,
    ) -> None:
        self.distance = _as_distance(distance)
        self._subjects_csr = None
        if sparse.issparse(subjects_factors) and self.distance != Distance.DOT:
            raise ValueError("To use `sparse.csr_matrix` distance must be `Distance.DOT`")  # rank_implicit.py:66-67
        if engine is None and self.distance != Distance.EUCLIDEAN and _is_cuda_tensor(objects_factors):
            # embeddings that already live on the GPU (transformer scorers keep `item_embs` on the device,
            # rectools/models/nn/transformers/lightning.py:391, :398): hand the device pointers over, no host round trip
            self._init_from_device_tensors(subjects_factors, objects_factors, tc_mode)
            return
        objects = _dense_f32(objects_factors)
        if sparse.issparse(subjects_factors):
            # EASE: the subjects are the user x item interaction CSR (rectools/models/ease.py:134-161).  The reference keeps the
            # matrix sparse and densifies only the requested rows (rank_implicit.py:236, :157-160); here the rows stay sparse
            # all the way into the SpMM scorer of the engine.
            csr = subjects_factors.tocsr()
            if csr.shape[1] != objects.shape[1]:
                raise ValueError("subject and object factors must have the same number of columns")
            self._subjects_csr = csr.astype(np.float32)
            self.n_subjects, self.n_objects = csr.shape[0], objects.shape[0]
            self.subjects_norms = self.subjects_dots = None
            self.engine = engine or Engine(objects, cosine=False, device=device, tc_mode=tc_mode)
            self._subjects, self._subjects_key = None, None
            self.last_stats = {}
            return
        subjects = _dense_f32(subjects_factors)
        if subjects.shape[1] != objects.shape[1]:
            raise ValueError("subject and object factors must have the same number of columns")
        self.n_subjects, self.n_objects = subjects.shape[0], objects.shape[0]
        subjects, objects, self.subjects_norms, self.subjects_dots = prepare_factors(self.distance, subjects, objects)
        self.engine = engine or Engine(objects, cosine=self.distance == Distance.COSINE, device=device, tc_mode=tc_mode)
        self._subjects, self._subjects_key = subjects, subjects_key
        self.engine.set_subjects(subjects, key=subjects_key, owner=self)
        self.last_stats: tp.Dict[str, tp.Any] = {}

    def _init_from_device_tensors(self, subjects_factors: tp.Any, objects_factors: tp.Any, tc_mode: str) -> None:
        import torch

        # fp16 / bf16 embeddings go to the engine as they are (widened exactly on the device: b200_rank_create_ex)
        dtypes = {torch.float32: _lib.DT_F32, torch.float16: _lib.DT_F16, torch.bfloat16: _lib.DT_BF16}
        objects = objects_factors.detach()
        if objects.dtype not in dtypes:
            objects = objects.to(torch.float32)
        objects = objects.contiguous()
        dev = objects.device
        subjects = subjects_factors
        if sparse.issparse(subjects):
            raise ValueError("CSR subjects need host object factors")
        if not hasattr(subjects, "detach"):
            subjects = torch.from_numpy(_dense_f32(subjects))
        subjects = subjects.detach().to(device=dev, dtype=torch.float32).contiguous()
        if subjects.shape[1] != objects.shape[1]:
            raise ValueError("subject and object factors must have the same number of columns")
        self.n_subjects, self.n_objects = int(subjects.shape[0]), int(objects.shape[0])
        self.subjects_norms = self.subjects_dots = None
        if self.distance == Distance.COSINE:
            norms = torch.linalg.vector_norm(subjects.double(), dim=1).float()
            norms[norms == 0] = 1e-10
            self.subjects_norms = norms.cpu().numpy()
        torch.cuda.current_stream(dev).synchronize()
        self._device_tensors = (subjects, objects)  # the engine references this memory: keep it alive
        self.engine = Engine(
            None, cosine=self.distance == Distance.COSINE, device=dev.index or 0, tc_mode=tc_mode,
            objects_device_ptr=objects.data_ptr(), shape=(self.n_objects, int(objects.shape[1])), objects_dtype=dtypes[objects.dtype],
        )
        self.engine.set_subjects_device(subjects.data_ptr(), self.n_subjects)
        self._subjects = self._subjects_key = None
        self.last_stats = {}

    # ------------------------------------------------------------------------------------------------------------
    def rank_padded(
        self,
        subject_ids: InternalIds,
        k: tp.Optional[int] = None,
        filter_pairs_csr: tp.Optional[sparse.csr_matrix] = None,
        sorted_object_whitelist: tp.Optional[np.ndarray] = None,
        flags: int = 0,
    ) -> tp.Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
        """`rank` without the ragged flattening: `(subject_ids, ids [n,k], scores [n,k], counts [n])`."""
        subject_ids = np.asarray(subject_ids, dtype=np.int64).reshape(-1)
        if filter_pairs_csr is not None and filter_pairs_csr.shape[0] != len(subject_ids):
            raise ValueError("Number of rows in `filter_pairs_csr` must be equal to `len(sublect_ids)`")
        if len(subject_ids) and (subject_ids.min() < 0 or subject_ids.max() >= self.n_subjects):
            raise IndexError("subject id out of range")
        whitelist = None
        n_pos = self.n_objects
        if sorted_object_whitelist is not None:
            whitelist = np.asarray(sorted_object_whitelist, dtype=np.int64).reshape(-1)
            check_whitelist(whitelist, self.n_objects)
            n_pos = len(whitelist)
        if k is None:
            k = n_pos  # rank_implicit.py:233-234
        if k <= 0:
            raise ValueError("`k` must be positive")
        indptr = indices = None
        if filter_pairs_csr is not None:
            csr = filter_pairs_csr if sparse.isspmatrix_csr(filter_pairs_csr) else sparse.csr_matrix(filter_pairs_csr)
            if not csr.has_sorted_indices:
                csr = csr.sorted_indices()
            indptr, indices = csr.indptr, csr.indices
        if n_pos == 0 or len(subject_ids) == 0:
            z = np.empty((len(subject_ids), 0))
            return subject_ids, z.astype(np.int32), z.astype(np.float32), np.zeros(len(subject_ids), np.int32)
        if self._subjects_csr is not None:
            rows = self._subjects_csr[subject_ids]  # CSR row gather: cheap, stays sparse (rank_implicit.py:236)
            ids, scores, counts = self.engine.topk(
                k, sparse_subjects=rows, indptr=indptr, indices=indices, whitelist=whitelist, flags=flags & ~_lib.Q_FORCE_TC
            )
            self.last_stats = self.engine.last_stats
            return subject_ids, ids, scores, counts
        if getattr(self, "_subjects", None) is not None and self.engine.subjects_owner is not self:
            # another ranker sharing this (cached) engine made its own subject factors resident in the meantime
            self.engine.set_subjects(self._subjects, key=self._subjects_key, owner=self)
        ids, scores, counts = self.engine.topk(
            k, subject_ids=subject_ids, indptr=indptr, indices=indices, whitelist=whitelist, flags=flags
        )
        self.last_stats = self.engine.last_stats
        return subject_ids, ids, scores, counts

    def rank(
        self,
        subject_ids: InternalIds,
        k: tp.Optional[int] = None,
        filter_pairs_csr: tp.Optional[sparse.csr_matrix] = None,
        sorted_object_whitelist: tp.Optional[np.ndarray] = None,
    ) -> tp.Tuple[InternalIds, InternalIds, Scores]:
        """Same contract as `ImplicitRanker.rank` (rank_implicit.py:187-280): flat `(subject ids repeated, object ids,
        scores)`, grouped by subject in input order, best first, filtered objects never returned."""
        subject_ids, ids, scores, counts = self.rank_padded(subject_ids, k, filter_pairs_csr, sorted_object_whitelist)
        all_subjects, all_ids, all_scores = flatten_padded(subject_ids, ids, scores, counts)
        if self.distance == Distance.COSINE:
            all_scores = all_scores / self.subjects_norms[all_subjects]  # rank_implicit.py:132-134
        elif self.distance == Distance.EUCLIDEAN:
            d2 = self.subjects_dots[all_subjects] - all_scores  # rank_implicit.py:136-140
            all_scores = np.sqrt(np.maximum(d2, 0)).astype(np.float32)
        return all_subjects, all_ids, all_scores
