"""Builders of UNMODIFIED reference models on synthetic data (shared by the CPU and GPU integration tests and bench.py's
`model_recommend` leg).  Needs the reference package on sys.path (`oracle.stage_reference.add_to_path()`)."""
import numpy as np


def synthetic_dataset(n_users, n_items, per_user, seed=0, external_offset=True):
    """`Dataset` over ~`per_user` distinct interactions per user.  Internal ids are 0..n-1 in order; external ids are shifted
    (users * 7 + 1000, items * 3 + 5) unless `external_offset=False` (pattern: tests/models/test_implicit_als.py:72-90)."""
    import pandas as pd
    from rectools import Columns
    from rectools.dataset import Dataset, IdMap, Interactions

    rng = np.random.default_rng(seed)
    users = np.repeat(np.arange(n_users, dtype=np.int64), per_user)
    items = rng.integers(0, n_items, size=n_users * per_user, dtype=np.int64)
    df = pd.DataFrame({Columns.User: users, Columns.Item: items})
    df = df.drop_duplicates([Columns.User, Columns.Item], ignore_index=True)
    df[Columns.Weight] = np.float64(1.0)
    df[Columns.Datetime] = pd.Timestamp("2024-01-01")
    user_ext = np.arange(n_users, dtype=np.int64) * 7 + 1000 if external_offset else np.arange(n_users, dtype=np.int64)
    item_ext = np.arange(n_items, dtype=np.int64) * 3 + 5 if external_offset else np.arange(n_items, dtype=np.int64)
    return Dataset(IdMap(user_ext), IdMap(item_ext), Interactions(df))


def injected_als(user_factors, item_factors):
    """`ImplicitALSWrapperModel` around a pre-"fitted" implicit ALS object carrying the given factors -- the injection of
    the reference's own test (tests/models/test_implicit_als.py:193-197); the stub's `AlternatingLeastSquares` is an
    attribute carrier (oracle/implicit_stub/implicit/cpu/als.py)."""
    from implicit.cpu.als import AlternatingLeastSquares
    from rectools.models import ImplicitALSWrapperModel

    base = AlternatingLeastSquares(factors=user_factors.shape[1], num_threads=0, iterations=1, random_state=0)
    base.user_factors = np.ascontiguousarray(user_factors, dtype=np.float32)
    base.item_factors = np.ascontiguousarray(item_factors, dtype=np.float32)
    wrapped = ImplicitALSWrapperModel(model=base, fit_features_together=False)
    wrapped.is_fitted = True
    wrapped.model = wrapped._model  # pylint: disable=protected-access
    return wrapped
