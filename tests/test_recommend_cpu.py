"""CPU: the vectorised `recommend()` returns the same table as the unmodified reference
`ModelBase.recommend` (rectools/models/base.py:385-519).  The reference runs here through `oracle/implicit_stub`; the
B200 ranker is replaced by the oracle-backed stand-in (`tests/helpers.OracleRanker`), so only the host logic is compared.
Needs the reference package: the checkout (build container) or its staged copy oracle/_ref (GPU box)."""
import sys

import numpy as np
import pytest

from oracle import stage_reference

pytestmark = pytest.mark.skipif(not stage_reference.available(), reason="reference package neither staged nor checked out")


@pytest.fixture(scope="module")
def fitted():
    added = stage_reference.add_to_path()
    import pandas as pd
    from rectools import Columns
    from rectools.dataset import Dataset
    from rectools.models import PureSVDModel

    rng = np.random.default_rng(0)
    n_users, n_items, n_inter = 300, 120, 6000
    df = pd.DataFrame(
        {
            Columns.User: rng.integers(0, n_users, n_inter) * 7 + 1000,  # external ids != internal ids
            Columns.Item: rng.integers(0, n_items, n_inter) * 3 + 5,  # (string ids trip the reference itself under pandas 3)
            Columns.Weight: 1.0,
            Columns.Datetime: pd.Timestamp("2024-01-01"),
        }
    ).drop_duplicates([Columns.User, Columns.Item])
    dataset = Dataset.construct(df)
    model = PureSVDModel(factors=8, random_state=0).fit(dataset)
    yield model, dataset, df
    stage_reference.remove_from_path(added)


def _same(ref, got):
    import pandas as pd

    pd.testing.assert_frame_equal(ref.reset_index(drop=True), got.reset_index(drop=True), check_exact=False, rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("filter_viewed", [True, False])
@pytest.mark.parametrize("add_rank_col", [True, False])
def test_all_users_match_reference(fitted, filter_viewed, add_rank_col):
    from rectools_b200.recommend import recommend
    from tests.helpers import OracleRanker

    model, dataset, _ = fitted
    users = dataset.user_id_map.external_ids
    ref = model.recommend(users, dataset, k=7, filter_viewed=filter_viewed, add_rank_col=add_rank_col)
    got = recommend(model, users, dataset, 7, filter_viewed, add_rank_col=add_rank_col, ranker_factory=OracleRanker)
    assert list(ref.columns) == list(got.columns) and [str(t) for t in ref.dtypes] == [str(t) for t in got.dtypes]
    _same(ref, got)


def test_user_subset_whitelist_and_ragged_rows(fitted):
    from rectools_b200.recommend import recommend
    from tests.helpers import OracleRanker

    model, dataset, df = fitted
    rng = np.random.default_rng(1)
    users = rng.permutation(dataset.user_id_map.external_ids)[:57]
    # a whitelist smaller than k plus the viewed filter: users get fewer than k rows (rank_implicit.py:107-118)
    items = df["item_id"].value_counts().index[:4].to_numpy()
    ref = model.recommend(users, dataset, k=6, filter_viewed=True, items_to_recommend=items)
    got = recommend(model, users, dataset, 6, True, items_to_recommend=items, ranker_factory=OracleRanker)
    assert len(ref) < 57 * 4 + 1 and ref.groupby("user_id").size().min() < 4
    _same(ref, got)
    # second call: the viewed-items CSR comes from the cache
    from rectools_b200 import recommend as rmod

    assert id(dataset.interactions.df) in sys.modules[rmod.__module__]._CSR_CACHE  # pylint: disable=protected-access
    _same(ref, recommend(model, users, dataset, 6, True, items_to_recommend=items, ranker_factory=OracleRanker))


def test_cold_targets_are_delegated(fitted):
    from rectools_b200.recommend import recommend
    from tests.helpers import OracleRanker

    model, dataset, _ = fitted
    users = np.concatenate([dataset.user_id_map.external_ids[:5], [10**9]])
    with pytest.raises(ValueError):
        recommend(model, users, dataset, 3, True, ranker_factory=OracleRanker)
    ref = model.recommend(users, dataset, k=3, filter_viewed=True, on_unsupported_targets="ignore")
    got = recommend(model, users, dataset, 3, True, on_unsupported_targets="ignore", ranker_factory=OracleRanker)
    _same(ref, got)


def test_install_patches_vector_model_recommend(fitted):
    import rectools.models.vector as vector
    from rectools.models.base import ModelBase

    import rectools_b200

    rectools_b200.install(fast_recommend=True)
    try:
        assert "recommend" in vector.VectorModel.__dict__ and "recommend_to_items" in vector.VectorModel.__dict__
    finally:
        rectools_b200.uninstall()
    assert "recommend" not in vector.VectorModel.__dict__ and vector.VectorModel.recommend is ModelBase.recommend
    assert vector.VectorModel.recommend_to_items is ModelBase.recommend_to_items


@pytest.mark.parametrize("filter_itself", [True, False])
@pytest.mark.parametrize("with_whitelist", [False, True])
def test_recommend_to_items_matches_reference(fitted, filter_itself, with_whitelist):
    """i2i = the same ranker with item vectors as subjects, k + 1 results, self-filter on the padded arrays."""
    from rectools_b200.recommend import recommend_to_items
    from tests.helpers import OracleRanker

    model, dataset, df = fitted
    rng = np.random.default_rng(2)
    targets = rng.permutation(dataset.item_id_map.external_ids)[:40]
    wl = None
    if with_whitelist:  # small enough that some targets get fewer than k rows; contains some of the targets themselves
        wl = np.concatenate([targets[:3], df["item_id"].value_counts().index[:4].to_numpy()])
    ref = model.recommend_to_items(targets, dataset, k=5, filter_itself=filter_itself, items_to_recommend=wl)
    got = recommend_to_items(model, targets, dataset, 5, filter_itself, items_to_recommend=wl, ranker_factory=OracleRanker)
    assert list(ref.columns) == list(got.columns) and [str(t) for t in ref.dtypes] == [str(t) for t in got.dtypes]
    _same(ref, got)


def test_recommend_to_items_repeated_targets_are_delegated(fitted):
    from rectools_b200.recommend import recommend_to_items
    from tests.helpers import OracleRanker

    model, dataset, _ = fitted
    t = dataset.item_id_map.external_ids[:3]
    targets = np.concatenate([t, t[:1]])
    ref = model.recommend_to_items(targets, dataset, k=4)
    _same(ref, recommend_to_items(model, targets, dataset, 4, ranker_factory=OracleRanker))


def test_repeated_users_are_delegated(fitted):
    """with repeated target users the reference's rank column runs across the repeats (`groupby(user).cumcount()`,
    base.py:778-791): the vectorised path hands such calls to the reference method."""
    from rectools_b200.recommend import recommend
    from tests.helpers import OracleRanker

    model, dataset, _ = fitted
    u = dataset.user_id_map.external_ids[:3]
    users = np.array([u[0], u[1], u[0]])
    ref = model.recommend(users, dataset, k=3, filter_viewed=True)
    got = recommend(model, users, dataset, 3, True, ranker_factory=OracleRanker)
    _same(ref, got)
    assert got["rank"].max() == 6


def test_viewed_csr_cache_notices_in_place_edits(fitted):
    """the cached viewed-items CSR is stamped with a digest of the (user, item) columns."""
    from rectools_b200.recommend import viewed_csr

    _, dataset, _ = fitted
    a = viewed_csr(dataset)
    assert viewed_csr(dataset) is a
    df = dataset.interactions.df
    old = df.loc[df.index[0], "item_id"]
    new = (old + 1) % dataset.item_id_map.size
    try:
        df.loc[df.index[0], "item_id"] = new
        b = viewed_csr(dataset)
        assert b is not a and (b != a).nnz > 0
    finally:
        df.loc[df.index[0], "item_id"] = old
    assert (viewed_csr(dataset) != a).nnz == 0


def test_warm_users_are_still_told_apart_with_the_cached_hot_count(fitted):
    """`n_hot_users` is answered from the stamped CSR cache entry (not a scan of the table per call): a user known only from
    the feature table is warm exactly as for the reference (base.py:676-700) -- refused, or dropped with a warning."""
    import pandas as pd
    from rectools.dataset import Dataset
    from rectools.models import PureSVDModel

    from rectools_b200.recommend import _viewed_entry, recommend
    from tests.helpers import OracleRanker

    _, _, df = fitted
    users = np.unique(df["user_id"].values)
    warm_id = int(users.max()) + 7
    feats = pd.DataFrame({"id": np.append(users, warm_id), "feature": "f", "value": 1.0})
    dataset = Dataset.construct(df, user_features_df=feats)
    model = PureSVDModel(factors=8, random_state=0).fit(dataset)
    assert dataset.user_id_map.size == len(users) + 1
    assert _viewed_entry(dataset)[1] == dataset.n_hot_users == len(users)
    ref = model.recommend(users[:50], dataset, k=5, filter_viewed=True)
    _same(ref, recommend(model, users[:50], dataset, 5, True, ranker_factory=OracleRanker))
    targets = np.append(users[:5], warm_id)
    with pytest.raises(ValueError, match="warm"):
        model.recommend(targets, dataset, k=5, filter_viewed=True)
    with pytest.raises(ValueError, match="warm"):
        recommend(model, targets, dataset, 5, True, ranker_factory=OracleRanker)
    with pytest.warns(UserWarning):
        ref = model.recommend(targets, dataset, k=5, filter_viewed=True, on_unsupported_targets="warn")
    with pytest.warns(UserWarning):
        got = recommend(model, targets, dataset, 5, True, on_unsupported_targets="warn", ranker_factory=OracleRanker)
    _same(ref, got)
    assert warm_id not in set(got["user_id"])


def test_threaded_table_columns_match_reference(fitted, monkeypatch):
    """Above 2^20 output rows the id gather / repeat / rank columns are written by row blocks on a thread pool: force that
    path at test size and compare with the reference table (u2i with unfilled slots, u2i full, i2i)."""
    import importlib

    from tests.helpers import OracleRanker

    rec = importlib.import_module("rectools_b200.recommend")  # (the package attribute of that name is the function)
    monkeypatch.setattr(rec, "_PAR_MIN", 1)
    model, dataset, _ = fitted
    users = dataset.user_id_map.external_ids
    for k, add_rank_col in ((7, True), (7, False), (dataset.item_id_map.size, True)):  # the last: ragged rows (-1 slots)
        ref = model.recommend(users, dataset, k=k, filter_viewed=True, add_rank_col=add_rank_col)
        _same(ref, rec.recommend(model, users, dataset, k, True, add_rank_col=add_rank_col, ranker_factory=OracleRanker))
    items = dataset.item_id_map.external_ids[:40]
    ref = model.recommend_to_items(items, dataset, k=6)
    _same(ref, rec.recommend_to_items(model, items, dataset, 6, ranker_factory=OracleRanker))
    table = np.arange(10, dtype=np.int64) * 3
    ids = np.array([[1, -1], [9, 0]], dtype=np.int32)
    assert rec.external_ids_of(table, ids, np.int64).tolist() == [[3, 0], [27, 0]]
    assert rec._repeat_rows(np.array([5, 6, 7]), 2).tolist() == [5, 5, 6, 6, 7, 7]
    assert rec._tile_rows(np.array([1, 2]), 3).tolist() == [1, 2, 1, 2, 1, 2]
