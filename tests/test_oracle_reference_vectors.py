"""CPU: known-answer vectors transcribed from the reference's own tests (values only), run through the oracle.

Sources (relative to /root/reference): tests/models/rank/test_rank.py:52-64 (fixture), :66-127 (plain),
:129-185 (filter_pairs_csr), :187-233 (whitelist); tests/models/rank/test_rank_implicit.py:51-104 (sentinel rules).
"""
import numpy as np
import pytest
from scipy import sparse

from oracle.topk_oracle import implicit_topk, neginf_score, rank_oracle

SUBJECTS = np.array([[-4, 0, 3], [0, 1, 2]])
OBJECTS = np.array([[-4, 0, 3], [0, 2, 4], [1, 10, 100]])


@pytest.mark.parametrize(
    "distance, expected_recs, expected_scores",
    (
        ("dot", [2, 0, 1, 2, 1, 0], [296, 25, 12, 210, 10, 6]),
        ("cosine", [0, 2, 1, 1, 2, 0], [1, 0.5890328, 0.5366563, 1, 0.9344414, 0.5366563]),
        ("euclidean", [0, 1, 2, 1, 0, 2], [0, 4.58257569, 97.64220399, 2.23606798, 4.24264069, 98.41747812]),
    ),
)
@pytest.mark.parametrize("accum", ["f32", "f64"])
def test_rank_known_answers(distance, expected_recs, expected_scores, accum):
    _, recs, scores = rank_oracle(distance, SUBJECTS, OBJECTS, [0, 1], k=3, accum=accum)
    np.testing.assert_equal(recs, expected_recs)
    np.testing.assert_almost_equal(scores, expected_scores, decimal=5)


@pytest.mark.parametrize(
    "distance, expected_recs, expected_scores",
    (
        ("dot", [2, 1, 2, 0], [296, 12, 210, 6]),
        ("cosine", [2, 1, 2, 0], [0.5890328, 0.5366563, 0.9344414, 0.5366563]),
        ("euclidean", [1, 2, 0, 2], [4.58257569, 97.64220399, 4.24264069, 98.41747812]),
    ),
)
def test_rank_with_filter_known_answers(distance, expected_recs, expected_scores):
    ui_csr = sparse.csr_matrix([[1, 0, 0], [0, 1, 0]])
    _, recs, scores = rank_oracle(distance, SUBJECTS, OBJECTS, [0, 1], k=3, filter_pairs_csr=ui_csr)
    np.testing.assert_equal(recs, expected_recs)
    np.testing.assert_almost_equal(scores, expected_scores, decimal=5)


@pytest.mark.parametrize(
    "distance, expected_recs, expected_scores",
    (
        ("dot", [2, 0, 2, 0], [296, 25, 210, 6]),
        ("cosine", [0, 2, 2, 0], [1, 0.5890328, 0.9344414, 0.5366563]),
        ("euclidean", [0, 2, 0, 2], [0, 97.64220399, 4.24264069, 98.41747812]),
    ),
)
def test_rank_with_whitelist_known_answers(distance, expected_recs, expected_scores):
    _, recs, scores = rank_oracle(distance, SUBJECTS, OBJECTS, [0, 1], k=3, sorted_object_whitelist=np.array([0, 2]))
    np.testing.assert_equal(recs, expected_recs)
    np.testing.assert_almost_equal(scores, expected_scores, decimal=5)


def test_shape_mismatch_raises():
    with pytest.raises(ValueError):
        rank_oracle("dot", SUBJECTS, OBJECTS, [0, 1], k=3, filter_pairs_csr=sparse.csr_matrix([[1, 0, 0]]))


def test_sparse_subjects_need_dot():
    with pytest.raises(ValueError):
        rank_oracle("cosine", sparse.csr_matrix(SUBJECTS), OBJECTS, [0, 1], k=3)


def test_neginf_sentinel_contract():
    dummy = np.array([[1, 2]], dtype=np.float32)
    neginf = implicit_topk(items=dummy, query=dummy, k=1, filter_items=np.array([0]))[1][0][0]
    assert neginf <= neginf_score() <= -1e38
