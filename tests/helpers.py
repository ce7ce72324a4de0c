"""Shared helpers for the parity tests (CPU and GPU)."""
import os

import numpy as np
from scipy import sparse

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_rank_case(case: int):
    inp = np.load(os.path.join(GOLDEN, f"rank_inputs_{case}.npz"))
    out = np.load(os.path.join(GOLDEN, f"rank_outputs_{case}.npz"))
    shape = tuple(int(x) for x in inp["csr_shape"])
    csr = sparse.csr_matrix(
        (np.ones(len(inp["csr_indices"]), dtype=np.float32), inp["csr_indices"], inp["csr_indptr"]), shape=shape
    )
    return inp, out, csr


def golden_keys(out):
    return sorted({k.rsplit("|", 1)[0] for k in out.files})


def parse_key(key: str):
    impl, dist, k, f, w = key.split("|")
    k = int(k[1:])
    return impl, dist, (None if k < 0 else k), f == "f1", w == "w1"


def synth_factors(n_users, n_items, d, seed=0):
    """SURVEY §8(d) synthetic inputs: standard_normal / sqrt(d), fp32, fixed seeds."""
    u = (np.random.default_rng(seed).standard_normal((n_users, d), dtype=np.float32) / np.sqrt(d)).astype(np.float32)
    i = (np.random.default_rng(seed + 1).standard_normal((n_items, d), dtype=np.float32) / np.sqrt(d)).astype(np.float32)
    return u, i


def synth_viewed_csr(n_users, n_items, per_user, seed=2):
    """~`per_user` viewed items per user (uniform draws, sorted within a row; rare duplicates are kept -- the filter
    is a set-membership test so they are harmless), int32 indices / int64 indptr, data = ones."""
    rng = np.random.default_rng(seed)
    cols = rng.integers(0, n_items, size=(n_users, per_user), dtype=np.int32)
    cols.sort(axis=1)
    indptr = np.arange(n_users + 1, dtype=np.int64) * per_user
    return sparse.csr_matrix((np.ones(cols.size, np.float32), cols.reshape(-1), indptr), shape=(n_users, n_items))


def ragged_to_padded(subjects, ids, scores, subject_ids, k):
    """Flat (subject, id, score) triplet -> [n, k] id / score arrays + counts (order preserved)."""
    n = len(subject_ids)
    out_ids = np.full((n, k), -1, dtype=np.int64)
    out_sc = np.full((n, k), np.nan, dtype=np.float32)
    counts = np.zeros(n, dtype=np.int64)
    pos = 0
    subjects = np.asarray(subjects)
    for r, sid in enumerate(subject_ids):
        c = 0
        while pos < len(subjects) and subjects[pos] == sid and c < k:
            out_ids[r, c] = ids[pos]
            out_sc[r, c] = scores[pos]
            pos += 1
            c += 1
        counts[r] = c
    return out_ids, out_sc, counts


def assert_same_ranking(ids, scores, ref_ids, ref_scores, rtol=2e-5, atol=2e-6, tie_tol=None, msg=""):
    """ids must be identical; where they differ the two rankings must be a permutation inside a near-tie window
    (`tie_tol`, relative to the score scale) -- the reference leaves tie order undefined (pure_svd.py:78-80)."""
    ids, ref_ids = np.asarray(ids), np.asarray(ref_ids)
    scores, ref_scores = np.asarray(scores, dtype=np.float64), np.asarray(ref_scores, dtype=np.float64)
    assert ids.shape == ref_ids.shape, f"{msg}: shape {ids.shape} vs {ref_ids.shape}"
    np.testing.assert_allclose(scores, ref_scores, rtol=max(rtol, tie_tol or 0), atol=atol, err_msg=msg)
    bad = np.nonzero(ids != ref_ids)[0]
    if len(bad) == 0:
        return 0
    assert tie_tol is not None, f"{msg}: {len(bad)} id mismatches, first at {bad[:5]}"
    scale = max(1e-30, float(np.abs(ref_scores).max()))
    for p in bad:
        # the item the reference put here must sit at a neighbouring position with an (almost) equal score
        assert abs(scores[p] - ref_scores[p]) <= tie_tol * scale, f"{msg}: non-tie mismatch at {p}"
    return len(bad)
