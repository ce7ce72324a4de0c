"""CPU: host logic of `B200Ranker` / `B200ImplicitRanker` around the C ABI, with the shared library replaced by a recording
stand-in (no compute): argument validation in the reference's terms, which calls reach the library, and when the resident
subject factors are (re-)uploaded for rankers that share one cached engine."""
import ctypes as C

import numpy as np
import pytest
from scipy import sparse


class RecordingLib:
    def __init__(self):
        self.calls = []

    def b200_rank_create_ex(self, out, ptr, dtype, n, d, dist, device, tc, flags):
        out._obj.value = 1 + len([c for c in self.calls if c[0] == "create"])  # pylint: disable=protected-access
        self.calls.append(("create", int(n), int(d), int(dist)))
        return 0

    def b200_rank_destroy(self, h):
        self.calls.append(("destroy", h.value))
        return 0

    def b200_rank_set_subjects(self, h, ptr, n, on_device):
        self.calls.append(("set_subjects", h.value, int(n)))
        return 0

    def b200_rank_set_id_offset(self, h, off):
        return 0

    def b200_rank_topk(self, h, q, stats):
        query = q._obj  # pylint: disable=protected-access
        self.calls.append(("topk", h.value, int(query.n_rows), int(query.k), bool(query.csr_indptr), bool(query.whitelist)))
        n, k = int(query.n_rows), int(query.k)
        C.memset(query.out_counts, 0, 4 * n)  # every row: no results
        return 0

    def b200_rank_last_error(self):
        return b""


@pytest.fixture()
def lib(monkeypatch):
    from rectools_b200 import _lib, integration

    rec = RecordingLib()
    monkeypatch.setattr(_lib, "_LIB", rec)
    integration.clear_engine_cache()
    yield rec
    integration._ENGINE_CACHE.clear()  # pylint: disable=protected-access


def test_reference_error_contract(lib):
    from rectools_b200 import B200Ranker

    u, i = np.ones((5, 3), np.float32), np.ones((7, 3), np.float32)
    ranker = B200Ranker("dot", u, i)
    with pytest.raises(ValueError, match="filter_pairs_csr"):  # rank_implicit.py:215-217
        ranker.rank([0, 1], k=2, filter_pairs_csr=sparse.csr_matrix((3, 7), dtype=np.float32))
    with pytest.raises(ValueError):  # rank_implicit.py:66-67
        B200Ranker("cosine", sparse.csr_matrix(u), i)
    with pytest.raises(ValueError):
        B200Ranker("dot", np.ones((5, 4), np.float32), i)
    with pytest.raises(IndexError):
        ranker.rank([0, 9], k=2)
    s, ids, sc = ranker.rank([], k=3)
    assert len(s) == len(ids) == len(sc) == 0 and not [c for c in lib.calls if c[0] == "topk"]
    s, ids, sc = ranker.rank([1, 0], k=None, sorted_object_whitelist=np.array([2, 5]))  # k=None -> all (whitelisted) objects
    assert [c for c in lib.calls if c[0] == "topk"][-1][2:] == (2, 2, False, True)
    assert len(ids) == 0  # the stand-in returns no rows; counts = 0 must flatten to nothing


def test_shared_cached_engine_keeps_the_right_subjects_resident(lib):
    from rectools_b200 import B200ImplicitRanker

    items = np.random.default_rng(0).random((50, 4), dtype=np.float32)
    users_a = np.random.default_rng(1).random((20, 4), dtype=np.float32)
    users_b = np.random.default_rng(2).random((30, 4), dtype=np.float32)
    r1 = B200ImplicitRanker("dot", users_a, items)
    r2 = B200ImplicitRanker("dot", users_a, items)  # same matrices: engine reused, subjects stay resident
    assert [c[0] for c in lib.calls] == ["create", "set_subjects"]
    r3 = B200ImplicitRanker("dot", users_b, items)  # same items, other subjects: one upload
    assert [c[0] for c in lib.calls] == ["create", "set_subjects", "set_subjects"] and lib.calls[-1][2] == 30
    r3.rank([0, 1], k=2)
    assert [c[0] for c in lib.calls][-1] == "topk"
    r1.rank([0, 1], k=2)  # r1's subjects were displaced by r3: uploaded again before ranking
    assert [c[0] for c in lib.calls][-2:] == ["set_subjects", "topk"] and lib.calls[-2][2] == 20
    r2.rank([0], k=1)  # same matrix as r1: nothing to upload
    assert [c[0] for c in lib.calls][-2:] == ["topk", "topk"]
    # COSINE needs its own engine (pre-normalised object copy)
    B200ImplicitRanker("cosine", users_a, items)
    assert [c for c in lib.calls if c[0] == "create"][-1][3] == 1 and len([c for c in lib.calls if c[0] == "create"]) == 2


def test_in_place_refit_is_detected(lib):
    """The engine / subject caches are keyed by the CONTENT of the factor matrices: an in-place change of a single element
    (a refit that reuses the arrays) must reach the device, an unchanged matrix must not be uploaded again."""
    from rectools_b200 import B200ImplicitRanker
    from rectools_b200.integration import content_hash

    items = np.random.default_rng(0).random((5000, 64), dtype=np.float32)
    users = np.random.default_rng(1).random((300, 64), dtype=np.float32)
    B200ImplicitRanker("dot", users, items)
    B200ImplicitRanker("dot", users, items)
    assert [c[0] for c in lib.calls] == ["create", "set_subjects"]
    items[4321, 17] += 1e-3  # not on any sampling stride
    B200ImplicitRanker("dot", users, items)
    assert [c[0] for c in lib.calls].count("create") == 2
    users[299, 63] = 0.5
    B200ImplicitRanker("dot", users, items)
    assert [c[0] for c in lib.calls].count("create") == 2 and [c[0] for c in lib.calls].count("set_subjects") == 3
    # position sensitive (two rows swapped), shape sensitive, tail bytes covered
    a = np.arange(3 * 100003, dtype=np.float32).reshape(3, 100003)
    b = a.copy()
    b[[0, 1]] = b[[1, 0]]
    assert content_hash(a) != content_hash(b) and content_hash(a) == content_hash(a.copy())
    assert content_hash(a) != content_hash(a.reshape(100003, 3))
    c = a.copy()
    c[-1, -1] += 1
    assert content_hash(a) != content_hash(c)


def test_whitelist_must_be_sorted_unique_and_in_range(lib):
    """the kernels merge viewed ids against ascending whitelist positions -- an unsorted whitelist is refused."""
    from rectools_b200 import B200Ranker

    ranker = B200Ranker("dot", np.ones((5, 3), np.float32), np.ones((7, 3), np.float32))
    ranker.rank([0], k=2, sorted_object_whitelist=np.array([1, 4, 6]))
    for bad in ([4, 1, 6], [1, 1, 4]):
        with pytest.raises(ValueError, match="sorted"):
            ranker.rank([0], k=2, sorted_object_whitelist=np.array(bad))
    for bad in ([1, 7], [-1, 2]):
        with pytest.raises(IndexError):
            ranker.rank([0], k=2, sorted_object_whitelist=np.array(bad))


def test_sparse_subjects_stay_sparse(lib):
    """EASE: CSR subject factors are never densified as a whole; the requested rows go to the library as CSR."""
    from rectools_b200 import B200Ranker

    x = sparse.random(50, 7, density=0.3, format="csr", dtype=np.float32, random_state=0)
    ranker = B200Ranker("dot", x, np.ones((7, 7), np.float32))
    assert [c[0] for c in lib.calls] == ["create"]  # no resident dense subjects
    ranker.rank([3, 1], k=2)
    assert lib.calls[-1][0] == "topk" and lib.calls[-1][2] == 2
    with pytest.raises(ValueError):
        B200Ranker("dot", sparse.csr_matrix((4, 6), dtype=np.float32), np.ones((7, 7), np.float32))
