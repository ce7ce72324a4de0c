"""Wiring into an installed RecTools (all imports of `rectools` are lazy: the package works without it).

Seams:
  * `VectorModel._recommend_u2i/_recommend_i2i` construct `ImplicitRanker(...)` inline (rectools/models/vector.py:66-72,
    :90-96; name imported at vector.py:28) and `EASEModel._recommend_u2i` does the same (rectools/models/ease.py:144,
    import at ease.py:31)  ->  `install()` rebinds that module-level name to `B200ImplicitRanker`.
  * transformer models take `similarity_module_type` (rectools/models/nn/transformers/base.py:219, :423); their
    `DistanceSimilarityModule._recommend_u2i` builds a `TorchRanker` (rectools/models/nn/transformers/similarity.py:127-132)
    ->  `make_similarity_module()` returns a subclass that builds a `B200TorchRanker` instead.
"""
from __future__ import annotations

import hashlib
import os
import threading
import typing as tp
from concurrent.futures import ThreadPoolExecutor

import numpy as np
from scipy import sparse

from .ranker import B200Ranker, Distance, Engine, _as_distance, _dense_f32

_ENGINE_CACHE: "tp.Dict[tp.Tuple, Engine]" = {}
_ENGINE_CACHE_MAX = 2
_CACHE_LOCK = threading.Lock()

_HASH_BLOCK = 1 << 15  # 64-bit words per block (256 KiB)
_HASH_WEIGHTS = np.random.default_rng(0x5EED).integers(1, 2**63, size=_HASH_BLOCK, dtype=np.uint64) | np.uint64(1)
_HASH_POOL: tp.Optional[ThreadPoolExecutor] = None


def content_hash(a: np.ndarray) -> bytes:
    """Digest of the WHOLE buffer of a C-contiguous array (shape and dtype included), position sensitive: every 64-bit
    word is multiplied by a fixed odd weight of its position inside a 256 KiB block and summed (wrapping), the block sums
    go through blake2b in order.  numpy releases the GIL, so the blocks are spread over a thread pool: ~10 ms per
    512 MB on a many-core host -- cheap next to the upload it saves, and unlike a sampled fingerprint it cannot miss an
    in-place refit."""
    global _HASH_POOL  # pylint: disable=global-statement
    a = np.ascontiguousarray(a)
    raw = a.reshape(-1).view(np.uint8)
    n_words = raw.size // 8
    words = raw[: n_words * 8].view(np.uint64)
    seg = 64 * _HASH_BLOCK  # words per task (16 MiB)

    def work(i: int) -> np.ndarray:
        chunk = words[i * seg : (i + 1) * seg]
        full = (len(chunk) // _HASH_BLOCK) * _HASH_BLOCK
        out = []
        if full:
            # (einsum: the weighted sums without the product temporary, 3x the multiply-then-sum rate; same wrapping result)
            out.append(np.einsum("ij,j->i", chunk[:full].reshape(-1, _HASH_BLOCK), _HASH_WEIGHTS))
        if full < len(chunk):
            rest = chunk[full:]
            out.append(np.array([np.einsum("i,i->", rest, _HASH_WEIGHTS[: len(rest)])], dtype=np.uint64))
        return np.concatenate(out) if out else np.empty(0, np.uint64)

    n_tasks = -(-n_words // seg) if n_words else 0
    if n_tasks > 1:
        if _HASH_POOL is None:
            _HASH_POOL = ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1), thread_name_prefix="b200hash")
        parts = list(_HASH_POOL.map(work, range(n_tasks)))
    else:
        parts = [work(i) for i in range(n_tasks)]
    h = hashlib.blake2b(digest_size=16)
    for part in parts:
        h.update(part.tobytes())
    h.update(raw[n_words * 8 :].tobytes())
    h.update(repr((a.shape, a.dtype.str)).encode())
    return h.digest()


def cached_engine(objects: np.ndarray, cosine: bool, device: int, tc_mode: str) -> Engine:
    """`VectorModel` builds a new ranker on every `recommend()` call (vector.py:66); keep the resident object factors
    across calls instead of re-uploading them (the reference GPU path re-uploads per call, rank_implicit.py:156).
    Keyed by the CONTENT of the matrix.  Evicted engines are only dropped from the cache: a ranker that still holds one
    keeps it alive, the device memory is released when the last reference goes (`Engine.__del__`)."""
    key = (content_hash(objects), cosine, device, tc_mode)
    with _CACHE_LOCK:
        eng = _ENGINE_CACHE.get(key)
        if eng is not None:
            return eng
    eng = Engine(objects, cosine=cosine, device=device, tc_mode=tc_mode)
    with _CACHE_LOCK:
        while len(_ENGINE_CACHE) >= _ENGINE_CACHE_MAX:
            _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE)))
        _ENGINE_CACHE[key] = eng
    return eng


def clear_engine_cache() -> None:
    with _CACHE_LOCK:
        _ENGINE_CACHE.clear()


class B200ImplicitRanker(B200Ranker):
    """`ImplicitRanker(distance, subjects_factors, objects_factors, num_threads=0, use_gpu=False)`-compatible
    constructor (rank_implicit.py:58-65) with a per-process engine cache keyed by the object matrix."""

    default_device: int = 0
    default_tc_mode: str = "auto"

    def __init__(self, distance, subjects_factors, objects_factors, num_threads: int = 0, use_gpu: bool = False) -> None:
        dist = _as_distance(distance)
        engine = None
        subjects_key = None
        if dist != Distance.EUCLIDEAN and isinstance(objects_factors, np.ndarray):
            objects = _dense_f32(objects_factors)
            engine = cached_engine(objects, dist == Distance.COSINE, self.default_device, self.default_tc_mode)
            objects_factors = objects
            if isinstance(subjects_factors, np.ndarray) and not sparse.issparse(subjects_factors):
                subjects_factors = _dense_f32(subjects_factors)
                subjects_key = content_hash(subjects_factors)  # same content as in the previous call: stays resident
        super().__init__(
            dist, subjects_factors, objects_factors, num_threads=num_threads, use_gpu=use_gpu,
            device=self.default_device, tc_mode=self.default_tc_mode, engine=engine, subjects_key=subjects_key,
        )


class B200TorchRanker(B200Ranker):
    """`TorchRanker(distance, device, subjects_factors, objects_factors, batch_size=128, dtype=torch.float32)`-compatible
    constructor (rectools/models/rank/rank_torch.py:59-67).  `batch_size` is meaningless here (no score matrix is ever
    materialised) and `dtype` other than fp32 is ignored: inputs are cast to fp32 like `_normalize_tensor` does by default.

    Difference kept from the reference: `TorchRanker` filters on CSR *values* != 0 (rank_torch.py:143) whereas the
    implicit path uses the stored structure; explicit zeros are dropped here to keep the torch semantics."""

    def __init__(self, distance, device, subjects_factors, objects_factors, batch_size: int = 128, dtype=None) -> None:
        dev_index = 0
        dev = str(device)
        if dev.startswith("cuda") and ":" in dev:
            dev_index = int(dev.split(":")[1])
        if hasattr(objects_factors, "detach") and dev.startswith("cuda") and not objects_factors.is_cuda:
            objects_factors = objects_factors.to(device)  # `TorchRanker` scores on `device` (rank_torch.py:135)
        super().__init__(distance, subjects_factors, objects_factors, device=dev_index)
        self.batch_size = batch_size

    def rank(self, subject_ids, k=None, filter_pairs_csr=None, sorted_object_whitelist=None):
        if filter_pairs_csr is not None and filter_pairs_csr.nnz and (filter_pairs_csr.data == 0).any():
            filter_pairs_csr = filter_pairs_csr.copy()
            filter_pairs_csr.eliminate_zeros()
        return super().rank(subject_ids, k, filter_pairs_csr, sorted_object_whitelist)


_ORIGINALS: tp.Dict[str, tp.Any] = {}
_FAST_KEY = "VectorModel.recommend"


def install(device: int = 0, tc_mode: str = "auto", fast_recommend: bool = True) -> None:
    """Route `VectorModel` (ALS / PureSVD / LightFM / BPR / DSSM) and `EASEModel` ranking through the B200 engine.

    `fast_recommend`: also give `VectorModel` the vectorised `recommend()` of `rectools_b200.recommend` (cached viewed-items
    CSR, id maps by array indexing, no per-user Python loop); warm / cold targets and context models still go through
    `ModelBase.recommend` (rectools/models/base.py:385-519)."""
    import importlib

    B200ImplicitRanker.default_device = device
    B200ImplicitRanker.default_tc_mode = tc_mode
    for modname in ("rectools.models.vector", "rectools.models.ease"):
        mod = importlib.import_module(modname)
        if modname not in _ORIGINALS:
            _ORIGINALS[modname] = mod.ImplicitRanker
        mod.ImplicitRanker = B200ImplicitRanker
    if fast_recommend and _FAST_KEY not in _ORIGINALS:
        from rectools.models.base import ModelBase
        from rectools.models.vector import VectorModel

        from .recommend import recommend as fast
        from .recommend import recommend_to_items as fast_i2i

        def _recommend(self, users, dataset, k, filter_viewed, items_to_recommend=None, add_rank_col=True,
                       on_unsupported_targets="raise", context=None):
            return fast(self, users, dataset, k, filter_viewed, items_to_recommend, add_rank_col, on_unsupported_targets, context,
                        reference_recommend=lambda *a, **kw: ModelBase.recommend(self, *a, **kw))

        def _recommend_to_items(self, target_items, dataset, k, filter_itself=True, items_to_recommend=None, add_rank_col=True,
                                on_unsupported_targets="raise"):
            return fast_i2i(self, target_items, dataset, k, filter_itself, items_to_recommend, add_rank_col, on_unsupported_targets,
                            reference_recommend=lambda *a, **kw: ModelBase.recommend_to_items(self, *a, **kw))

        _recommend.__doc__ = ModelBase.recommend.__doc__
        _recommend_to_items.__doc__ = ModelBase.recommend_to_items.__doc__
        # (None: the attribute is inherited from ModelBase)
        _ORIGINALS[_FAST_KEY] = (VectorModel.__dict__.get("recommend"), VectorModel.__dict__.get("recommend_to_items"))
        VectorModel.recommend = _recommend
        VectorModel.recommend_to_items = _recommend_to_items


def uninstall() -> None:
    import importlib

    if _FAST_KEY in _ORIGINALS:
        from rectools.models.vector import VectorModel

        for name, orig in zip(("recommend", "recommend_to_items"), _ORIGINALS.pop(_FAST_KEY)):
            if orig is None:
                delattr(VectorModel, name)
            else:
                setattr(VectorModel, name, orig)
    for modname, orig in list(_ORIGINALS.items()):
        importlib.import_module(modname).ImplicitRanker = orig
        del _ORIGINALS[modname]
    clear_engine_cache()


def make_similarity_module(ranker_factory: tp.Optional[tp.Callable[..., tp.Any]] = None) -> type:
    """`similarity_module_type` for SASRec / BERT4Rec / HSTU (rectools/models/nn/transformers/base.py:219, :423): the
    reference's `DistanceSimilarityModule` with `B200TorchRanker` as the scorer of `_recommend_u2i`
    (similarity.py:117-140).  `item_embs` stays on its device (and in its dtype: fp16 / bf16 embeddings are handed to the
    engine as they are); the filter stays a CSR (the reference densifies [batch, n_items] per batch, rank_torch.py:138-144).
    `ranker_factory`: another `TorchRanker`-signature class (the CPU tests plug an oracle-backed stand-in in)."""
    from rectools.models.nn.transformers.similarity import DistanceSimilarityModule  # needs torch only

    factory = ranker_factory or B200TorchRanker

    class B200DistanceSimilarityModule(DistanceSimilarityModule):
        def _recommend_u2i(self, user_embs, item_embs, user_ids, k, sorted_item_ids_to_recommend, ui_csr_for_filter):
            ranker = factory(
                distance=self.distance, device=item_embs.device, subjects_factors=user_embs[user_ids],
                objects_factors=item_embs,
            )
            user_ids_indices, all_reco_ids, all_scores = ranker.rank(
                subject_ids=np.arange(len(user_ids)), k=k, filter_pairs_csr=ui_csr_for_filter,
                sorted_object_whitelist=sorted_item_ids_to_recommend,
            )
            return user_ids[user_ids_indices], all_reco_ids, all_scores

    return B200DistanceSimilarityModule
