"""Numpy restatement of the reference ranking path (test infrastructure only).

Two layers, each citing the reference lines it follows (paths relative to
/root/reference):

* :func:`implicit_topk` restates the third-party ``implicit.cpu.topk.topk``
  as it is called at ``rectools/models/rank/rank_implicit.py:264-272``:
  fp32 ``query @ items.T``, optional division by ``item_norms``, ``-FLT_MAX`` at
  the stored entries of ``filter_query_items`` and at ``filter_items``
  (sentinel contract: ``tests/models/rank/test_rank_implicit.py:51-71``),
  per-row top-k sorted by score descending.
* :func:`rank_oracle` restates ``ImplicitRanker.__init__`` + ``.rank`` +
  ``_process_implicit_scores`` (``rank_implicit.py:58-81, 187-280, 83-146``).

``accum="f32"`` follows the reference arithmetic (fp32 BLAS product).
``accum="f64"`` is the adjudicator the CUDA path is specified against: the dot
product is accumulated in float64 and rounded once to fp32, so the ranking does
not depend on a BLAS summation order.  Ties: (score desc, id asc) -- the
reference leaves tie order implementation-defined (``pure_svd.py:78-80``).
"""

from __future__ import annotations

import typing as tp

import numpy as np
from scipy import sparse

FLT_MAX = np.finfo(np.float32).max
NEG_SENTINEL = np.float32(-FLT_MAX)


def neginf_score() -> float:
    """``ImplicitRanker._get_neginf_score`` (rank_implicit.py:83-92): bits(-FLT_MAX) - 1."""
    return float(np.asarray(np.asarray(-FLT_MAX, dtype=np.float32).view(np.uint32) - 1, dtype=np.uint32).view(np.float32))


def calc_norms(factors: np.ndarray, accum: str = "f32") -> np.ndarray:
    """``ImplicitRanker._calc_norms(avoid_zeros=True)`` (rank_implicit.py:98-105)."""
    if accum == "f64":
        norms = np.sqrt((factors.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)
    else:
        norms = np.linalg.norm(factors, axis=1)
    norms = np.asarray(norms, dtype=np.float32).copy()
    norms[norms == 0] = 1e-10
    return norms


def _select_topk_rows(scores: np.ndarray, k: int) -> tp.Tuple[np.ndarray, np.ndarray]:
    """Row-wise top-k of a dense fp32 score block, ordered (score desc, id asc)."""
    n = scores.shape[1]
    k = min(k, n)
    if k < n:
        # candidates: everything >= the k-th largest value (keeps all ties at the cut)
        kth = np.partition(scores, n - k, axis=1)[:, n - k]
    ids_out = np.empty((scores.shape[0], k), dtype=np.int32)
    sc_out = np.empty((scores.shape[0], k), dtype=np.float32)
    for r in range(scores.shape[0]):
        row = scores[r]
        if k < n:
            cand = np.nonzero(row >= kth[r])[0]
        else:
            cand = np.arange(n)
        # lexsort: last key is primary -> (-score asc, id asc)
        order = np.lexsort((cand, -row[cand].astype(np.float64)))[:k]
        sel = cand[order]
        ids_out[r] = sel
        sc_out[r] = row[sel]
    return ids_out, sc_out


def _score_block(query: np.ndarray, items: np.ndarray, item_norms: tp.Optional[np.ndarray], accum: str) -> np.ndarray:
    if accum == "f64":  # (`items` arrives already widened: implicit_topk converts the catalogue once, not per batch)
        s = query.astype(np.float64) @ items.T
        if item_norms is not None:
            s = s / item_norms.astype(np.float64)[None, :]
        return s.astype(np.float32)
    s = query @ items.T  # fp32 BLAS sgemm, as numpy/implicit do
    if item_norms is not None:
        s = s / item_norms[None, :]
    return s.astype(np.float32, copy=False)


def implicit_topk(
    items: np.ndarray,
    query: np.ndarray,
    k: int,
    item_norms: tp.Optional[np.ndarray] = None,
    filter_query_items: tp.Optional[sparse.csr_matrix] = None,
    filter_items: tp.Optional[np.ndarray] = None,
    num_threads: int = 0,  # pylint: disable=unused-argument
    accum: str = "f32",
    batch: int = 512,
) -> tp.Tuple[np.ndarray, np.ndarray]:
    """Restatement of ``implicit.cpu.topk.topk`` (call site rank_implicit.py:264-272).

    Returns ``(ids int32 [Q,k], scores fp32 [Q,k])``; masked entries carry
    ``-FLT_MAX`` and may occupy the tail when fewer than ``k`` items survive.
    """
    items = np.ascontiguousarray(items, dtype=np.float32)
    if sparse.issparse(query):
        query = np.asarray(query.todense())
    query = np.ascontiguousarray(query, dtype=np.float32)
    n_q, n_items = query.shape[0], items.shape[0]
    k = min(int(k), n_items)
    if item_norms is not None:
        item_norms = np.asarray(item_norms, dtype=np.float32).reshape(-1)
    if filter_query_items is not None:
        filter_query_items = sparse.csr_matrix(filter_query_items)
        indptr, indices = filter_query_items.indptr, filter_query_items.indices
    ids = np.empty((n_q, k), dtype=np.int32)
    scores = np.empty((n_q, k), dtype=np.float32)
    items_acc = items.astype(np.float64) if accum == "f64" else items
    for start in range(0, n_q, batch):
        stop = min(start + batch, n_q)
        s = _score_block(query[start:stop], items_acc, item_norms, accum)
        if filter_query_items is not None:
            for r in range(start, stop):
                cols = indices[indptr[r] : indptr[r + 1]]
                cols = cols[cols < n_items]
                s[r - start, cols] = NEG_SENTINEL
        if filter_items is not None:
            s[:, np.asarray(filter_items, dtype=np.int64)] = NEG_SENTINEL
        ids[start:stop], scores[start:stop] = _select_topk_rows(s, k)
    return ids, scores


def rank_oracle(  # pylint: disable=too-many-locals,too-many-branches
    distance: str,
    subjects_factors: tp.Union[np.ndarray, sparse.csr_matrix],
    objects_factors: np.ndarray,
    subject_ids: tp.Sequence[int],
    k: tp.Optional[int] = None,
    filter_pairs_csr: tp.Optional[sparse.csr_matrix] = None,
    sorted_object_whitelist: tp.Optional[np.ndarray] = None,
    accum: str = "f32",
    batch: int = 512,
) -> tp.Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Restatement of ``ImplicitRanker(distance, S, O).rank(...)``.  ``batch``: subjects per score block (memory bound of
    the restatement, no effect on the result).

    Steps follow rank_implicit.py: fp32 casts (:70-71), shape check (:215-217),
    whitelist gather and CSR column restriction (:219-226), subject gather
    (:236), COSINE object norms (:238-240), EUCLIDEAN augmentation (:242-246),
    ``real_k`` (:248), the top-k call (:264-272), whitelist remap (:274-275) and
    ``_process_implicit_scores`` (:120-146) incl. the trailing-sentinel strip
    (:107-118) and the COSINE / EUCLIDEAN post-scaling (:132-140).
    """
    distance = str(getattr(distance, "value", distance))
    subject_ids = np.asarray(subject_ids, dtype=np.int64)
    if sparse.issparse(subjects_factors):
        if distance != "dot":
            raise ValueError("To use `sparse.csr_matrix` distance must be `Distance.DOT`")
        subjects_factors = np.asarray(subjects_factors.todense())
    s_all = np.asarray(subjects_factors).astype(np.float32)
    o_all = np.asarray(objects_factors).astype(np.float32)

    if filter_pairs_csr is not None and filter_pairs_csr.shape[0] != len(subject_ids):
        raise ValueError("Number of rows in `filter_pairs_csr` must be equal to `len(sublect_ids)`")

    if sorted_object_whitelist is not None:
        wl = np.asarray(sorted_object_whitelist, dtype=np.int64)
        objects = o_all[wl]
        filt = None
        if filter_pairs_csr is not None:
            csr = sparse.csr_matrix(filter_pairs_csr)
            wl_in = wl[wl < csr.shape[1]]
            filt = sparse.csr_matrix(csr[:, wl_in])
            if wl_in.size < wl.size:  # CSR narrower than the catalogue: missing columns are unfiltered
                filt = sparse.csr_matrix((filt.data, filt.indices, filt.indptr), shape=(csr.shape[0], wl.size))
    else:
        wl = None
        objects = o_all
        filt = None if filter_pairs_csr is None else sparse.csr_matrix(filter_pairs_csr)

    if k is None:
        k = objects.shape[0]

    subjects = s_all[subject_ids]
    norms = None
    if distance == "cosine":
        norms = calc_norms(objects, accum)
        subj_norms = calc_norms(s_all, accum)
    elif distance == "euclidean":
        subj_dots = (s_all**2).sum(axis=1)
        # note the float64 promotion through np.ones / np.hstack in the reference (:244-245)
        subjects = np.hstack((-np.ones((subjects.shape[0], 1)), 2 * subjects)).astype(np.float32)
        objects = np.hstack(((objects**2).sum(axis=1).reshape(-1, 1), objects)).astype(np.float32)
    elif distance != "dot":
        raise ValueError(f"Unexpected distance `{distance}`")

    real_k = min(int(k), objects.shape[0])
    ids, scores = implicit_topk(objects, subjects, real_k, norms, filt, None, accum=accum, batch=batch)
    if wl is not None:
        ids = wl[ids]

    min_score = np.float32(neginf_score())
    out_subj, out_ids, out_scores = [], [], []
    for row, sid in enumerate(subject_ids):
        sc = scores[row]
        n_masked = 0
        for el in sc[::-1]:
            if el <= min_score:
                n_masked += 1
            else:
                break
        keep = len(sc) - n_masked
        rel = sc[:keep].copy()
        if distance == "cosine":
            rel /= subj_norms[sid]
        elif distance == "euclidean":
            rel = np.sqrt(np.maximum(subj_dots[sid] - rel, 0)).astype(np.float32)
        out_subj.append(np.full(keep, sid, dtype=np.int64))
        out_ids.append(np.asarray(ids[row][:keep], dtype=np.int64))
        out_scores.append(rel.astype(np.float32))
    if not out_subj:
        return np.empty(0, np.int64), np.empty(0, np.int64), np.empty(0, np.float32)
    return np.concatenate(out_subj), np.concatenate(out_ids), np.concatenate(out_scores)


def recommend_from_scores_numpy(scores: np.ndarray, k: int) -> tp.Tuple[np.ndarray, np.ndarray]:
    """The reference's numpy top-k idiom (rectools/models/utils.py:102-106), no lists."""
    n_reco = min(k, scores.size)
    pos = scores.argpartition(-n_reco)[-n_reco:]
    order = pos[scores[pos].argsort()[::-1]]
    return order, scores[order]
