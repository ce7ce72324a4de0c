"""CPU oracle for the RecTools vector-ranking hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and there only as the
checker (or as the timed CPU baseline), never as a fallback of the CUDA path.

Parity status: pinned against the reference's own known-answer vectors
(``tests/models/rank/test_rank.py:52-127`` and friends, transcribed in
``tests/test_oracle_reference_vectors.py``) and against golden fixtures under
``tests/golden/`` generated here by running the *unmodified* reference
(``TorchRanker``, ``ImplicitRanker`` through ``oracle/implicit_stub``,
``PureSVDModel``) -- see ``oracle/make_golden.py``.  The third-party arithmetic
(``implicit.cpu.topk.topk``, pm-implicit 0.7.3, not vendored, not installed)
is restated from its call site ``rectools/models/rank/rank_implicit.py:264-272``
and the sentinel contract ``tests/models/rank/test_rank_implicit.py:51-71``.
Tie order and exact fp32 summation order are NOT pinned by the reference
(``rectools/models/pure_svd.py:78-80``); the oracle fixes (score desc, id asc).
"""
