"""Vectorised `recommend()` around the B200 ranker: SURVEY.md section 8(f) rank 1 (the callers either side of the hot path).

`ModelBase.recommend` (rectools/models/base.py:385-519) spends, around the ranker call,
  * `Dataset.get_user_item_matrix` -- a COO -> CSR rebuild of ALL interactions on every call (vector.py:58-60,
    dataset.py:314-348, interactions.py:162-175): ~1.3 s per 10^7 interactions;
  * a per-user Python loop over the ranker output (rank_implicit.py:120-146), removed by `B200Ranker.rank_padded`;
  * pandas `Series.reindex` for the internal -> external id maps (base.py:755-764, utils/indexing.py:124-132) and
    `groupby(sort=False).cumcount()` for the rank column (base.py:778-791): ~0.15 s per 10^6 rows.
`recommend()` below returns the same table (same columns, dtypes, row order, values) for the case every in-repo
caller of the vector path hits -- all target users hot -- with the viewed-items CSR cached per interactions table, id maps
applied by array indexing and the rank column written directly; anything else (warm / cold targets, models that need a
recommend context) is delegated to the model's own `recommend`, so behaviour is never narrower than the reference's.

Works on any object with the `VectorModel` surface (`_get_u2i_vectors`, `u2i_dist`, the `ModelBase` helper methods); the
only rectools import is lazy (column names), so the module imports without rectools installed.
"""
from __future__ import annotations

import typing as tp
import weakref

import numpy as np

from .ranker import Distance, _as_distance

USER_COL, ITEM_COL, SCORE_COL, RANK_COL = "user_id", "item_id", "score", "rank"  # rectools/columns.py:21-27
TARGET_ITEM_COL = "target_item_id"  # rectools/columns.py:23

_CSR_CACHE: "tp.Dict[int, tp.Tuple[tp.Any, tp.Any, tp.Any, tp.Optional[int]]]" = {}


def viewed_csr(dataset: tp.Any) -> tp.Any:
    """`dataset.get_user_item_matrix(include_weights=False)` (vector.py:59), built once per interactions table.

    The reference rebuilds this CSR from the interactions DataFrame on every `recommend()` call; it only depends on that
    (immutable by convention) table, so it is cached by the table's identity and dropped when the table is collected."""
    return _viewed_entry(dataset)[0]


def _viewed_entry(dataset: tp.Any) -> tp.Tuple[tp.Any, tp.Optional[int]]:
    """(the cached CSR, `dataset.n_hot_users` of the same stamped table or None when the table cannot be stamped)."""
    from .integration import content_hash

    df = dataset.interactions.df
    key = id(df)
    # the CSR structure depends on the (user, item) columns only: their content digest catches in-place edits of the table
    try:
        stamp = (len(df), content_hash(np.asarray(df[USER_COL].values)), content_hash(np.asarray(df[ITEM_COL].values)),
                 dataset.user_id_map.size, dataset.item_id_map.size)
    except (TypeError, KeyError, AttributeError):  # not a DataFrame (duck-typed datasets): identity of the table only
        stamp = None
    hit = _CSR_CACHE.get(key)
    if hit is not None and hit[0]() is df and hit[2] == stamp:
        return hit[1], hit[3]
    csr = dataset.get_user_item_matrix(include_weights=False)
    if not csr.has_sorted_indices:
        csr.sort_indices()
    # `Dataset.n_hot_users` (dataset.py:176-184) is a max over the user column on every access (0.06 s per 10^8 rows): the
    # stamp above covers that column, so the value is kept with the CSR
    n_hot = int(np.asarray(df[USER_COL].values).max()) + 1 if stamp is not None and len(df) > 0 else None
    try:
        ref = weakref.ref(df, lambda _r, key=key: _CSR_CACHE.pop(key, None))
    except TypeError:  # not weak-referenceable: do not cache
        return csr, n_hot
    _CSR_CACHE[key] = (ref, csr, stamp, n_hot)
    return csr, n_hot


class _KnownHotUsers:  # pylint: disable=too-few-public-methods
    """The dataset, with `n_hot_users` answered from the stamped cache entry instead of a scan of the interactions."""

    def __init__(self, dataset: tp.Any, n_hot_users: int) -> None:
        self._dataset = dataset
        self.n_hot_users = n_hot_users

    def __getattr__(self, name: str) -> tp.Any:
        return getattr(self._dataset, name)


def clear_viewed_cache() -> None:
    _CSR_CACHE.clear()


def _rows_of(csr: tp.Any, user_ids: np.ndarray) -> tp.Any:
    """`user_items[user_ids]` (vector.py:60) without the copy when the targets are all users in order."""
    n = csr.shape[0]
    if len(user_ids) == n and (n == 0 or (user_ids[0] == 0 and user_ids[-1] == n - 1 and (np.diff(user_ids) == 1).all())):
        return csr
    rows = csr[user_ids]
    rows.has_sorted_indices = True  # row selection keeps the (sorted) order inside every row: spare the ranker an O(nnz) check
    return rows


_PAR_MIN = 1 << 20  # output elements below which the table columns are written by the calling thread alone
_PAR_POOL: tp.Optional[tp.Any] = None


def _row_blocks(n_rows: int, work: tp.Callable[[int, int], None]) -> None:
    """`work(r0, r1)` over blocks of rows on a small thread pool (numpy copies and gathers release the GIL; at 10^7 output
    rows the single-threaded column writes cost as much as a quarter of the GPU pass)."""
    global _PAR_POOL  # pylint: disable=global-statement
    import os
    from concurrent.futures import ThreadPoolExecutor

    n_tasks = min(16, os.cpu_count() or 1)
    if _PAR_POOL is None:
        _PAR_POOL = ThreadPoolExecutor(max_workers=n_tasks, thread_name_prefix="b200table")
    step = -(-n_rows // n_tasks)
    list(_PAR_POOL.map(lambda i: work(i * step, min(n_rows, (i + 1) * step)), range(-(-n_rows // step))))


def external_ids_of(table: np.ndarray, ids: np.ndarray, dtype: tp.Any) -> np.ndarray:
    """`table[max(ids, 0)]` as `dtype` for a padded [n, k] id array (slots beyond a row's count hold -1 and are masked out
    later): internal -> external ids by array indexing (`IdMap.external_ids` is sorted by internal id, identifiers.py:124-126)."""
    table = np.asarray(table)
    if ids.size < _PAR_MIN or ids.ndim != 2 or table.dtype.hasobject or len(table) == 0:
        return np.asarray(table[np.maximum(ids, 0)], dtype=dtype)
    out = np.empty(ids.shape, dtype=table.dtype)

    def work(r0: int, r1: int) -> None:
        np.take(table, ids[r0:r1], mode="clip", out=out[r0:r1])  # clip: -1 -> 0, as the maximum above

    _row_blocks(ids.shape[0], work)
    return np.asarray(out, dtype=dtype)


def _repeat_rows(values: np.ndarray, k: int) -> np.ndarray:
    """`np.repeat(values, k)` / (values [k], tiled per row when `values` is the row pattern) written by row blocks."""
    n = len(values)
    if n * k < _PAR_MIN or values.dtype.hasobject:
        return np.repeat(values, k)
    out = np.empty((n, k), dtype=values.dtype)

    def work(r0: int, r1: int) -> None:
        out[r0:r1] = values[r0:r1, None]

    _row_blocks(n, work)
    return out.reshape(-1)


def _tile_rows(pattern: np.ndarray, n: int) -> np.ndarray:
    """`np.tile(pattern, n)` written by row blocks."""
    k = len(pattern)
    if n * k < _PAR_MIN:
        return np.tile(pattern, n)
    out = np.empty((n, k), dtype=pattern.dtype)

    def work(r0: int, r1: int) -> None:
        out[r0:r1] = pattern[None, :]

    _row_blocks(n, work)
    return out.reshape(-1)


def _has_repeats(ids: np.ndarray) -> bool:
    """Any id listed twice?  (Strictly ascending targets -- the usual "all users" call -- are decided by one linear pass.)"""
    if len(ids) < 2 or bool((np.diff(ids) > 0).all()):
        return False
    return len(np.unique(ids)) != len(ids)


def finalize_scores(ranker: tp.Any, subject_ids: np.ndarray, scores: np.ndarray) -> np.ndarray:
    """Distance post-scaling of `_process_implicit_scores` (rank_implicit.py:132-140) on the padded [n, k] array."""
    dist = _as_distance(ranker.distance)
    if dist == Distance.COSINE:
        return (scores / ranker.subjects_norms[subject_ids][:, None]).astype(np.float32, copy=False)
    if dist == Distance.EUCLIDEAN:
        d2 = ranker.subjects_dots[subject_ids][:, None] - scores
        return np.sqrt(np.maximum(d2, 0)).astype(np.float32)
    return scores


def reco_table(
    target_ext: np.ndarray,
    item_ext: np.ndarray,
    scores: np.ndarray,
    counts: np.ndarray,
    k_out: int,
    add_rank_col: bool,
    target_col: str = USER_COL,
    keep: tp.Optional[np.ndarray] = None,
) -> tp.Any:
    """`_make_reco_table` (base.py:778-791) from padded arrays: one row per returned pair, targets in input order.
    `keep` (bool [n, k_out]) selects the pairs explicitly (i2i: the target itself removed); default: the first `counts`."""
    import pandas as pd

    full = keep is None and k_out > 0 and len(counts) > 0 and int(counts.min()) == k_out
    if k_out == 0 or len(target_ext) == 0:
        targets, items, sc = target_ext[:0], item_ext.reshape(-1)[:0], scores.reshape(-1)[:0]
        ranks = np.empty(0, dtype=np.int64)
    elif full:
        targets, items, sc = _repeat_rows(target_ext, k_out), item_ext.reshape(-1), scores.reshape(-1)
        ranks = _tile_rows(np.arange(1, k_out + 1, dtype=np.int64), len(counts)) if add_rank_col else None
    else:
        mask = keep if keep is not None else np.arange(k_out, dtype=np.int32)[None, :] < counts[:, None]
        per_row = mask.sum(axis=1)
        targets, items, sc = np.repeat(target_ext, per_row), item_ext[mask], scores[mask]
        ranks = np.cumsum(mask, axis=1, dtype=np.int64)[mask]
    # copy=False: the columns are fresh arrays (or views of the ranker's own output) that nobody else holds -- pandas would
    # otherwise stack and copy them (0.5 s per 10^7 rows, as much as the whole GPU pass at U = 1M)
    cols = {target_col: targets, ITEM_COL: items, SCORE_COL: sc}
    if add_rank_col:
        cols[RANK_COL] = ranks
    return pd.DataFrame(cols, copy=False)


def recommend(  # pylint: disable=too-many-locals
    model: tp.Any,
    users: tp.Any,
    dataset: tp.Any,
    k: int,
    filter_viewed: bool,
    items_to_recommend: tp.Optional[tp.Any] = None,
    add_rank_col: bool = True,
    on_unsupported_targets: str = "raise",
    context: tp.Any = None,
    ranker_factory: tp.Optional[tp.Callable[..., tp.Any]] = None,
    reference_recommend: tp.Optional[tp.Callable[..., tp.Any]] = None,
) -> tp.Any:
    """Same contract as `ModelBase.recommend` (base.py:385-519) for `VectorModel`s; see the module docstring.

    `ranker_factory(distance, user_vectors, item_vectors)` defaults to `B200ImplicitRanker` (engine cached per item
    matrix); `reference_recommend` is the bound method to delegate to (default `model.recommend`)."""
    fallback = reference_recommend or model.recommend

    def delegate() -> tp.Any:
        return fallback(users, dataset, k, filter_viewed, items_to_recommend=items_to_recommend, add_rank_col=add_rank_col,
                        on_unsupported_targets=on_unsupported_targets, context=context)

    if context is not None or getattr(model, "require_recommend_context", False) or not hasattr(model, "_get_u2i_vectors"):
        return delegate()
    model._check_is_fitted()  # pylint: disable=protected-access
    model._check_k(k)  # pylint: disable=protected-access
    user_type = dataset.user_id_map.external_dtype
    item_type = dataset.item_id_map.external_dtype
    ds = model._custom_transform_dataset_u2i(dataset, users, on_unsupported_targets, None)  # pylint: disable=protected-access
    whitelist = model._get_sorted_item_ids_to_recommend(items_to_recommend, ds)  # pylint: disable=protected-access
    csr_all, n_hot = _viewed_entry(ds) if filter_viewed else (None, None)
    split_ds = ds if n_hot is None else _KnownHotUsers(ds, n_hot)
    hot, warm, cold = model._split_targets_by_hot_warm_cold(users, split_ds, "user")  # pylint: disable=protected-access
    hot, warm, cold = model._check_targets_are_valid(hot, warm, cold, "user", on_unsupported_targets)  # pylint: disable=protected-access
    hot = np.asarray(hot, dtype=np.int64)
    if np.size(warm) > 0 or np.size(cold) > 0 or _has_repeats(hot):
        # (repeated targets: the reference's rank column runs across the repeats, `groupby(user).cumcount()`, base.py:778-791)
        return delegate()

    csr = _rows_of(csr_all, hot) if (filter_viewed and hot.size) else None
    user_vectors, item_vectors = model._get_u2i_vectors(ds)  # pylint: disable=protected-access
    if ranker_factory is None:
        from .integration import B200ImplicitRanker

        ranker_factory = B200ImplicitRanker
    ranker = ranker_factory(model.u2i_dist, user_vectors, item_vectors)
    if hot.size:
        _, ids, scores, counts = ranker.rank_padded(hot, k, csr, whitelist)
        scores = finalize_scores(ranker, hot, scores)
    else:
        ids = np.empty((0, 0), dtype=np.int32)
        scores = np.empty((0, 0), dtype=np.float32)
        counts = np.empty(0, dtype=np.int32)
    k_out = ids.shape[1]
    # unfilled slots (id -1, beyond `counts`) are masked out in reco_table
    user_ext = np.asarray(ds.user_id_map.external_ids[hot], dtype=user_type)
    item_ext = external_ids_of(ds.item_id_map.external_ids, ids, item_type)
    return reco_table(user_ext, item_ext, np.asarray(scores, dtype=np.float32), counts, k_out, add_rank_col)


def recommend_to_items(  # pylint: disable=too-many-locals
    model: tp.Any,
    target_items: tp.Any,
    dataset: tp.Any,
    k: int,
    filter_itself: bool = True,
    items_to_recommend: tp.Optional[tp.Any] = None,
    add_rank_col: bool = True,
    on_unsupported_targets: str = "raise",
    ranker_factory: tp.Optional[tp.Callable[..., tp.Any]] = None,
    reference_recommend: tp.Optional[tp.Callable[..., tp.Any]] = None,
) -> tp.Any:
    """Same contract as `ModelBase.recommend_to_items` (base.py:521-646) for `VectorModel`s: the same
    kernel with subjects = item vectors (`_get_i2i_vectors`, `i2i_dist`), `k + 1` results when the target itself is filtered,
    and that filter (`_filter_item_itself_from_i2i_reco`, base.py:745-753: a DataFrame query + groupby.head) done on the
    padded arrays.  Warm / cold or repeated targets are delegated to the reference method."""
    fallback = reference_recommend or model.recommend_to_items

    def delegate() -> tp.Any:
        return fallback(target_items, dataset, k, filter_itself=filter_itself, items_to_recommend=items_to_recommend,
                        add_rank_col=add_rank_col, on_unsupported_targets=on_unsupported_targets)

    if not hasattr(model, "_get_i2i_vectors"):
        return delegate()
    model._check_is_fitted()  # pylint: disable=protected-access
    model._check_k(k)  # pylint: disable=protected-access
    item_type = dataset.item_id_map.external_dtype
    ds = model._custom_transform_dataset_i2i(dataset, target_items, on_unsupported_targets)  # pylint: disable=protected-access
    whitelist = model._get_sorted_item_ids_to_recommend(items_to_recommend, ds)  # pylint: disable=protected-access
    hot, warm, cold = model._split_targets_by_hot_warm_cold(target_items, ds, "item")  # pylint: disable=protected-access
    hot, warm, cold = model._check_targets_are_valid(hot, warm, cold, "item", on_unsupported_targets)  # pylint: disable=protected-access
    hot = np.asarray(hot, dtype=np.int64)
    if np.size(warm) > 0 or np.size(cold) > 0 or _has_repeats(hot):
        return delegate()  # (the reference groups the self-filter by target id: repeated targets share one group)

    requested_k = k + 1 if filter_itself else k  # base.py:603
    vectors_1, vectors_2 = model._get_i2i_vectors(ds)  # pylint: disable=protected-access
    if ranker_factory is None:
        from .integration import B200ImplicitRanker

        ranker_factory = B200ImplicitRanker
    ranker = ranker_factory(model.i2i_dist, vectors_1, vectors_2)
    if hot.size:
        _, ids, scores, counts = ranker.rank_padded(hot, requested_k, None, whitelist)
        scores = finalize_scores(ranker, hot, scores)
    else:
        ids, scores, counts = np.empty((0, 0), np.int32), np.empty((0, 0), np.float32), np.empty(0, np.int32)
    k_out = ids.shape[1]
    keep = np.arange(k_out, dtype=np.int32)[None, :] < counts[:, None]
    if filter_itself and k_out:
        keep &= ids != hot[:, None]
        keep &= np.cumsum(keep, axis=1) <= k  # the first k of what is left (groupby("tid").head(k))
    target_ext = np.asarray(ds.item_id_map.external_ids[hot], dtype=item_type)
    item_ext = external_ids_of(ds.item_id_map.external_ids, ids, item_type)
    return reco_table(target_ext, item_ext, np.asarray(scores, dtype=np.float32), keep.sum(axis=1), k_out, add_rank_col,
                      target_col=TARGET_ITEM_COL, keep=keep)
