"""Generate tests/golden/*.npz by running the UNMODIFIED reference here (test infrastructure only).

Run in the build container (it reads /root/reference, which does not exist on the GPU box):

    PYTHONPATH=/root/reference:oracle/implicit_stub python oracle/make_golden.py

What is recorded (inputs + the reference's outputs, fp32 / int64):
  * torch_*    -- `rectools.models.rank.TorchRanker` (rank_torch.py:77-177): an independent in-repo implementation
                  of the Ranker contract (torch CPU matmul + torch.topk), DOT and COSINE, with/without
                  `filter_pairs_csr`, with/without `sorted_object_whitelist`, k in {1, 10, None}.
  * implicit_* -- `rectools.models.rank.ImplicitRanker` (rank_implicit.py:187-280) driven through the `implicit` stub
                  whose `topk` is oracle/topk_oracle.py::implicit_topk (pins prologue/epilogue semantics, incl. the
                  < k rows case and EUCLIDEAN).
  * puresvd_c1 -- BASELINE config 1: `PureSVDModel(factors=32)` fit on synthetic 6 040 x 3 706 interactions
                  (MovieLens-1M shape), `recommend(k=10, filter_viewed=True)`; stores the fitted factor matrices,
                  the filter CSR the model builds (vector.py:58-60) and the returned (user, item, score) table.
Seeds are fixed; continuous random factors => no intra-user score ties.
"""

from __future__ import annotations

import os
import sys

import numpy as np
import pandas as pd
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")


def _random_case(seed: int, n_subj: int, n_obj: int, d: int, nnz_per_row: int):
    rng = np.random.default_rng(seed)
    s = (rng.standard_normal((n_subj, d)) / np.sqrt(d)).astype(np.float32)
    o = (rng.standard_normal((n_obj, d)) / np.sqrt(d)).astype(np.float32)
    subject_ids = rng.permutation(n_subj)[: max(1, n_subj * 3 // 4)].astype(np.int64)
    rows, cols = [], []
    for r in range(len(subject_ids)):
        m = int(rng.integers(0, nnz_per_row + 1))
        c = rng.choice(n_obj, size=min(m, n_obj), replace=False)
        rows.extend([r] * len(c))
        cols.extend(c.tolist())
    csr = sparse.csr_matrix(
        (np.ones(len(rows), dtype=np.float32), (rows, cols)), shape=(len(subject_ids), n_obj), dtype=np.float32
    )
    csr.sort_indices()
    whitelist = np.sort(rng.choice(n_obj, size=max(2, n_obj // 3), replace=False)).astype(np.int64)
    return s, o, subject_ids, csr, whitelist


def _save(name: str, **arrays) -> None:
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1024:.1f} KiB")


def ranker_cases() -> None:
    """One `rank_inputs_<c>.npz` per seeded case + one `rank_outputs_<c>.npz` holding every (impl, distance, k,
    filter, whitelist) combination under keys `<impl>|<distance>|k<k>|f<0/1>|w<0/1>|{subjects,ids,scores}`."""
    from rectools.models.rank import Distance, ImplicitRanker, TorchRanker

    for case, (n_subj, n_obj, d, nnz) in enumerate([(48, 300, 16, 12), (33, 1000, 40, 60), (17, 64, 7, 70)]):
        s, o, subject_ids, csr, whitelist = _random_case(100 + case, n_subj, n_obj, d, nnz)
        _save(
            f"rank_inputs_{case}",
            subjects=s,
            objects=o,
            subject_ids=subject_ids,
            csr_indptr=csr.indptr.astype(np.int64),
            csr_indices=csr.indices.astype(np.int32),
            csr_shape=np.asarray(csr.shape, dtype=np.int64),
            whitelist=whitelist,
        )
        outputs = {}
        for dist in (Distance.DOT, Distance.COSINE, Distance.EUCLIDEAN):
            rankers = (
                ("torch", TorchRanker(distance=dist, device="cpu", subjects_factors=s, objects_factors=o)),
                ("implicit", ImplicitRanker(dist, s, o)),
            )
            for k in (1, 10, 100) if case == 1 else (1, 10, None):
                for use_filter in (False, True):
                    for use_wl in (False, True):
                        for impl, ranker in rankers:
                            u, i, sc = ranker.rank(
                                subject_ids=subject_ids,
                                k=k,
                                filter_pairs_csr=csr if use_filter else None,
                                sorted_object_whitelist=whitelist if use_wl else None,
                            )
                            key = f"{impl}|{dist.value}|k{-1 if k is None else k}|f{int(use_filter)}|w{int(use_wl)}"
                            outputs[key + "|subjects"] = np.asarray(u, dtype=np.int64)
                            outputs[key + "|ids"] = np.asarray(i, dtype=np.int32)
                            outputs[key + "|scores"] = np.asarray(sc, dtype=np.float32)
        _save(f"rank_outputs_{case}", **outputs)


def puresvd_c1() -> None:
    from rectools import Columns
    from rectools.dataset import Dataset
    from rectools.models import PureSVDModel

    rng = np.random.default_rng(0)
    n_users, n_items, draws = 6040, 3706, 1_000_000
    users = rng.integers(0, n_users, size=draws)
    items = (rng.zipf(1.3, size=draws) - 1) % n_items
    df = pd.DataFrame({Columns.User: users, Columns.Item: items}).drop_duplicates()
    df[Columns.Weight] = 1.0
    df[Columns.Datetime] = pd.Timestamp("2024-01-01")
    dataset = Dataset.construct(df)
    model = PureSVDModel(factors=32, random_state=0).fit(dataset)

    ext_users = np.sort(dataset.user_id_map.external_ids)[:768]
    reco = model.recommend(users=ext_users, dataset=dataset, k=10, filter_viewed=True)
    reco_nf = model.recommend(users=ext_users, dataset=dataset, k=10, filter_viewed=False)

    user_vectors, item_vectors = model._get_u2i_vectors(dataset)  # pylint: disable=protected-access
    int_users = dataset.user_id_map.convert_to_internal(ext_users)
    ui = dataset.get_user_item_matrix(include_weights=False)[int_users]
    ui.sort_indices()

    def table(r):
        return dict(
            users=dataset.user_id_map.convert_to_internal(r[Columns.User].to_numpy()).astype(np.int64),
            items=dataset.item_id_map.convert_to_internal(r[Columns.Item].to_numpy()).astype(np.int64),
            scores=r[Columns.Score].to_numpy().astype(np.float32),
        )

    t, tn = table(reco), table(reco_nf)
    _save(
        "puresvd_c1",
        user_factors=user_vectors.astype(np.float32),
        item_factors=item_vectors.astype(np.float32),
        subject_ids=int_users.astype(np.int64),
        csr_indptr=ui.indptr.astype(np.int64),
        csr_indices=ui.indices.astype(np.int32),
        csr_shape=np.asarray(ui.shape, dtype=np.int64),
        out_subjects=t["users"],
        out_ids=t["items"],
        out_scores=t["scores"],
        out_nf_subjects=tn["users"],
        out_nf_ids=tn["items"],
        out_nf_scores=tn["scores"],
        n_interactions=np.asarray([len(df)], dtype=np.int64),
    )


if __name__ == "__main__":
    if not os.path.isdir("/root/reference/rectools"):
        sys.exit("make_golden.py needs the reference checkout at /root/reference (build container only)")
    ranker_cases()
    puresvd_c1()
