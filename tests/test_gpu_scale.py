"""GPU parity at the sizes BASELINE.json quotes: >= 4096 sampled subjects against the fp64 oracle at
N = 1M (DOT K = 10, COSINE K = 100) and N = 625 K, d = 256, bf16, K = 20 (config 5's per-GPU shard), plus a case that drives
rows through the second-chance pass AND the exhaustive re-rank at N = 1M, and the EASE shape (users x items CSR subjects,
items x items dense weights) at 100 K x 20 K.  The certificate failure rate depends on N, not on the number of subjects,
so a 4096-row batch against the full catalogue pins the same machinery the headline runs through."""
import numpy as np
import pytest
from scipy import sparse

from oracle.topk_oracle import rank_oracle
from tests.helpers import synth_viewed_csr

pytestmark = pytest.mark.gpu

BLOCK = 65536


def gen_factors(n, d, seed):
    """bench.py's generator (blocks of 64 K rows, seeded per block): N(0,1)/sqrt(d), fp32."""
    out = np.empty((n, d), dtype=np.float32)
    for b in range((n + BLOCK - 1) // BLOCK):
        r0, r1 = b * BLOCK, min((b + 1) * BLOCK, n)
        blk = np.random.default_rng([seed, b]).standard_normal((r1 - r0, d), dtype=np.float32)
        blk *= np.float32(1.0 / np.sqrt(d))
        out[r0:r1] = blk
    return out


def _check(ranker, distance, users, items, k, csr, ids, scores, counts, rtol=3e-7):
    n = users.shape[0]
    assert (counts == k).all()
    _, oid, osc = rank_oracle(distance, users, items, np.arange(n), k, csr, accum="f64")
    if distance == "cosine":
        un = np.sqrt(np.einsum("ij,ij->i", users, users, dtype=np.float64)).astype(np.float32)
        osc = osc * np.repeat(un, k)
    np.testing.assert_array_equal(ids.reshape(-1), oid, err_msg=str(ranker.last_stats))
    np.testing.assert_allclose(scores.reshape(-1), osc, rtol=rtol, atol=1e-9)


@pytest.fixture(scope="module")
def catalogue_1m():
    return gen_factors(1_000_000, 128, 1)


def test_c2_shape_dot_k10_n1m(rb, catalogue_1m, monkeypatch):
    from rectools_b200 import _lib

    monkeypatch.setenv("B200_TC_SPLITS", "1")  # two lists per row over the whole stream, as in a 1M-subject call

    items = catalogue_1m
    users = gen_factors(4096, 128, 0)
    csr = synth_viewed_csr(4096, items.shape[0], 100)
    eng = rb.Engine(items, cosine=False)
    ids, scores, counts = eng.topk(10, subjects=users, indptr=csr.indptr, indices=csr.indices, flags=_lib.Q_FORCE_TC)
    assert eng.last_stats["path"] == 1 and eng.last_stats["n_fallback_rows"] <= 40
    _check(eng, "dot", users, items, 10, csr, ids, scores, counts)
    eng.close()


def test_c3_shape_cosine_k100_n1m(rb, catalogue_1m):
    from rectools_b200 import _lib

    items = catalogue_1m
    users = gen_factors(4096, 128, 0)
    csr = synth_viewed_csr(4096, items.shape[0], 100)
    eng = rb.Engine(items, cosine=True)
    ids, scores, counts = eng.topk(100, subjects=users, indptr=csr.indptr, indices=csr.indices, flags=_lib.Q_FORCE_TC)
    st = eng.last_stats
    assert st["path"] == 1 and st["wide"] == 1 and st["n_fallback_rows"] <= 4096 // 10, st
    _check(eng, "cosine", users, items, 100, csr, ids, scores, counts)
    eng.close()


def test_c5_shard_shape_d256_bf16_k20(rb, monkeypatch):
    """Config 5's per-GPU shard: 625 K items, d = 256, bf16 item embeddings handed over as a DEVICE tensor, K = 20."""
    import torch

    monkeypatch.setenv("B200_TC_SPLITS", "1")
    n_items, d, k = 625_000, 256, 20
    items = torch.from_numpy(gen_factors(n_items, d, 1)).to(torch.bfloat16)
    users = torch.from_numpy(gen_factors(4096, d, 0)).to(torch.bfloat16).float()
    csr = synth_viewed_csr(4096, n_items, 50)
    ranker = rb.B200Ranker("dot", users, items.to("cuda:0"))
    _, ids, scores, counts = ranker.rank_padded(np.arange(4096), k, csr)
    st = ranker.last_stats
    assert st["path"] == 1 and st["tc_dtype"] == 2, st  # bf16 tensor-core candidates
    # the oracle sees the same bf16-rounded values
    _check(ranker, "dot", users.numpy(), items.float().numpy(), k, csr, ids, scores, counts)


def test_second_chance_and_exhaustive_rerank_at_n1m(rb, catalogue_1m, monkeypatch):
    """Planted near-ties at N = 1M.  A list only fails its certificate when ties SATURATE it: for 32 subjects 40 objects within
    ~1e-7 (relative) of each other sit on top -- ~20 per column half, more than the K' = 12 slots, fewer than the 32 of the
    second-chance pass; for 8 subjects 80 EXACT duplicates -- ~40 per half, more than 32 slots: only the exhaustive fp64
    re-rank can order them.  ids must equal the oracle's (score desc, id asc)."""
    from rectools_b200 import _lib

    monkeypatch.setenv("B200_TC_SPLITS", "1")  # (object splits would spread the planted ties over 2 x splits lists)
    items = catalogue_1m.copy()
    n_items = items.shape[0]
    users = gen_factors(4096, 128, 3)
    rng = np.random.default_rng(11)
    free = rng.permutation(n_items)[: 32 * 40 + 8 * 80]
    pos = 0
    for r in range(32):
        v = 1.5 * users[r] / np.linalg.norm(users[r])
        for j in range(40):
            items[free[pos]] = (v * np.float32(1.0 + 1.2e-7 * j)).astype(np.float32)
            pos += 1
    for r in range(100, 108):
        v = (1.5 * users[r] / np.linalg.norm(users[r])).astype(np.float32)
        items[free[pos : pos + 80]] = v
        pos += 80
    csr = synth_viewed_csr(4096, n_items, 100)
    eng = rb.Engine(items, cosine=False)
    ids, scores, counts = eng.topk(10, subjects=users, indptr=csr.indptr, indices=csr.indices, flags=_lib.Q_FORCE_TC)
    st = eng.last_stats
    assert st["path"] == 1 and st["n_fallback_rows"] >= 40 and st["n_exact_rows"] >= 8, st
    _check(eng, "dot", users, items, 10, csr, ids, scores, counts)
    eng.close()


def test_ease_shape_sparse_subjects_100k_x_20k(rb):
    """`EASEModel._recommend_u2i` (ease.py:134-161): subjects = users x items interaction CSR, objects = items x items dense
    weights (zero diagonal), DOT, filter = the same CSR.  100 K x 20 K: densifying the subjects would be 8 GB; the engine
    keeps them sparse (SpMM scorer).  Sampled rows against the fp64 oracle."""
    n_users, n_items, k = 100_000, 20_000, 10
    rng = np.random.default_rng(5)
    w = rng.standard_normal((n_items, n_items), dtype=np.float32) * np.float32(0.05)
    np.fill_diagonal(w, 0.0)
    cols = rng.integers(0, n_items, size=(n_users, 40), dtype=np.int32)
    cols.sort(axis=1)
    x = sparse.csr_matrix((np.ones(cols.size, np.float32), cols.reshape(-1), np.arange(n_users + 1, dtype=np.int64) * 40), shape=(n_users, n_items))
    x.sum_duplicates()
    ranker = rb.B200Ranker("dot", x, w)  # objects_factors = weight.T in the reference; any dense [n_items, n_items] matrix here
    sids = np.arange(n_users)
    _, ids, scores, counts = ranker.rank_padded(sids, k, x)
    assert ranker.last_stats["path"] == 2 and (counts == k).all()
    sel = np.arange(0, n_users, 49)[:2048]
    _, oid, osc = rank_oracle("dot", x[sel], w, np.arange(len(sel)), k, x[sel], accum="f64")
    np.testing.assert_array_equal(ids[sel].reshape(-1), oid)
    np.testing.assert_allclose(scores[sel].reshape(-1), osc, rtol=3e-7, atol=1e-9)
    # subset of subjects in another order + whitelist + k > 32 (two selection passes)
    sub = rng.permutation(n_users)[:3000]
    wl = np.arange(0, n_items, 3)
    s2, i2, sc2 = ranker.rank(sub, 40, x[sub], wl)
    _, oid2, osc2 = rank_oracle("dot", x[sub[:300]], w, np.arange(300), 40, x[sub[:300]], wl, accum="f64")
    n = len(oid2)
    np.testing.assert_array_equal(i2[:n], oid2)
    np.testing.assert_allclose(sc2[:n], osc2, rtol=3e-7, atol=1e-9)
