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

import typing as tp

import numpy as np

from .ranker import B200Ranker, Distance, Engine, _as_distance, _dense_f32

_ENGINE_CACHE: "tp.Dict[tp.Tuple, Engine]" = {}
_ENGINE_CACHE_MAX = 2


def _fingerprint(a: np.ndarray) -> tp.Tuple:
    """Cheap identity of a factor matrix: address, shape and a strided content sample (detects in-place refits)."""
    flat = a.reshape(-1)
    step = max(1, flat.size // 4096)
    sample = flat[::step][:4096]
    return (a.ctypes.data, a.shape, float(np.float64(sample.sum())), float(np.abs(sample).sum()))


def cached_engine(objects: np.ndarray, cosine: bool, device: int, tc_mode: str) -> Engine:
    """`VectorModel` builds a new ranker on every `recommend()` call (vector.py:66); keep the resident object factors
    across calls instead of re-uploading them (the reference GPU path re-uploads per call, rank_implicit.py:156)."""
    key = (_fingerprint(objects), cosine, device, tc_mode)
    eng = _ENGINE_CACHE.get(key)
    if eng is None:
        while len(_ENGINE_CACHE) >= _ENGINE_CACHE_MAX:
            _ENGINE_CACHE.pop(next(iter(_ENGINE_CACHE))).close()
        eng = Engine(objects, cosine=cosine, device=device, tc_mode=tc_mode)
        _ENGINE_CACHE[key] = eng
    return eng


def clear_engine_cache() -> None:
    while _ENGINE_CACHE:
        _ENGINE_CACHE.popitem()[1].close()


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
            if isinstance(subjects_factors, np.ndarray) and subjects_factors.dtype == np.float32 and subjects_factors.flags.c_contiguous:
                subjects_key = _fingerprint(subjects_factors)  # same matrix as in the previous call: stays resident
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


def make_similarity_module() -> type:
    """`similarity_module_type` for SASRec / BERT4Rec / HSTU: same module, `B200TorchRanker` as the scorer."""
    from rectools.models.nn.transformers.similarity import DistanceSimilarityModule  # needs torch + lightning

    class B200DistanceSimilarityModule(DistanceSimilarityModule):
        def _recommend_u2i(self, user_embs, item_embs, user_ids, k, sorted_item_ids_to_recommend, ui_csr_for_filter):
            ranker = B200TorchRanker(
                distance=self.distance, device=item_embs.device, subjects_factors=user_embs[user_ids],
                objects_factors=item_embs,
            )
            user_ids_indices, all_reco_ids, all_scores = ranker.rank(
                subject_ids=np.arange(len(user_ids)), k=k, filter_pairs_csr=ui_csr_for_filter,
                sorted_object_whitelist=sorted_item_ids_to_recommend,
            )
            return user_ids[user_ids_indices], all_reco_ids, all_scores

    return B200DistanceSimilarityModule

