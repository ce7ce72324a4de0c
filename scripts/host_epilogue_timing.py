#!/usr/bin/env python
"""CPU timing of the code AROUND the ranker: the reference's `ModelBase.recommend` (rebuilds the viewed-items CSR, maps ids with
pandas reindex, groupby-cumcount rank column) against `rectools_b200.recommend` (cached CSR, array indexing), both with the
same instantaneous fake ranker, so the difference is host logic only.  Needs the reference checkout + oracle/implicit_stub.

    python scripts/host_epilogue_timing.py [n_users] [n_items] [per_user]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [ROOT, "/root/reference", os.path.join(ROOT, "oracle", "implicit_stub")]
import pandas as pd  # noqa: E402
from rectools import Columns  # noqa: E402
from rectools.dataset import Dataset  # noqa: E402
from rectools.models import PureSVDModel  # noqa: E402
import rectools.models.vector as vector  # noqa: E402

from rectools_b200.recommend import recommend  # noqa: E402

n_users = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
n_items = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
per_user = int(sys.argv[3]) if len(sys.argv) > 3 else 50
K = 10
rng = np.random.default_rng(0)
df = pd.DataFrame({
    Columns.User: np.repeat(np.arange(n_users), per_user),
    Columns.Item: rng.integers(0, n_items, n_users * per_user),
    Columns.Weight: 1.0,
    Columns.Datetime: pd.Timestamp("2024-01-01"),
}).drop_duplicates([Columns.User, Columns.Item])
dataset = Dataset.construct(df)
model = PureSVDModel(factors=4).fit(dataset)
IDS = rng.integers(0, dataset.item_id_map.size, (n_users, K)).astype(np.int32)
SC = np.sort(rng.random((n_users, K), dtype=np.float32), axis=1)[:, ::-1].copy()


class FakeRanker:  # answers instantly: only the host code around it is timed
    def __init__(self, distance, u, i, num_threads=0, use_gpu=False):
        self.distance = "dot"

    def rank(self, subject_ids, k=None, filter_pairs_csr=None, sorted_object_whitelist=None):
        s = np.asarray(subject_ids)
        return np.repeat(s, K), IDS[s].reshape(-1).astype(np.int64), SC[s].reshape(-1)

    def rank_padded(self, subject_ids, k=None, filter_pairs_csr=None, sorted_object_whitelist=None, flags=0):
        s = np.asarray(subject_ids)
        return s, IDS[s], SC[s], np.full(len(s), K, np.int32)


vector.ImplicitRanker = FakeRanker
users = dataset.user_id_map.external_ids
for name, fn in (
    ("reference ModelBase.recommend", lambda: model.recommend(users, dataset, K, True)),
    ("rectools_b200.recommend (1st call: builds + caches the CSR)", lambda: recommend(model, users, dataset, K, True, ranker_factory=FakeRanker)),
    ("rectools_b200.recommend (cached CSR)", lambda: recommend(model, users, dataset, K, True, ranker_factory=FakeRanker)),
):
    t0 = time.perf_counter()
    out = fn()
    dt = time.perf_counter() - t0
    print(f"{name}: {dt:.3f} s for {n_users} users x K={K} ({len(df)} interactions) -> {n_users / dt:,.0f} users/s host-side ceiling, {len(out)} rows")
