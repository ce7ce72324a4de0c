"""CPU baseline = restatement of `implicit.cpu.topk.topk` at native speed (test infra / bench baseline only).

Mirrors the upstream structure named at the call site rectools/models/rank/rank_implicit.py:264-272:
per query batch a BLAS ``sgemm`` (numpy -> OpenBLAS, all host threads) followed by a native,
OpenMP-parallel per-row select (``oracle/select_ref.c``).  ``rank_cpu`` adds the reference's own
prologue/epilogue semantics in vectorised numpy so that a whole ``ImplicitRanker.rank`` call is timed.
"""

from __future__ import annotations

import ctypes
import os
import subprocess
import typing as tp

import numpy as np
from scipy import sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB: tp.Optional[ctypes.CDLL] = None


def build() -> str:
    out = os.path.join(_HERE, "_build", "libselect_ref.so")
    src = os.path.join(_HERE, "select_ref.c")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return out


def _lib() -> ctypes.CDLL:
    global _LIB  # pylint: disable=global-statement
    if _LIB is None:
        lib = ctypes.CDLL(build())
        lib.ref_topk_select.restype = ctypes.c_int
        lib.ref_topk_select.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32,
        ]
        lib.ref_num_threads.restype = ctypes.c_int
        _LIB = lib
    return _LIB


def num_threads() -> int:
    return int(_lib().ref_num_threads())


def use_all_threads() -> int:
    """Give BLAS (numpy's sgemm) and the OpenMP select every core this process may run on, whatever OMP_NUM_THREADS says
    (torchrun exports OMP_NUM_THREADS=1 to its workers: the round-1 reference arm silently ran on one thread at N > 1).
    Returns the thread count; pass it to `topk_cpu(num_threads=...)`."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits

        threadpool_limits(limits=n)  # process-wide, stays in force
    except Exception:  # pylint: disable=broad-except
        pass
    return n


def topk_cpu(
    items: np.ndarray,
    query: np.ndarray,
    k: int,
    item_norms: tp.Optional[np.ndarray] = None,
    filter_query_items: tp.Optional[sparse.csr_matrix] = None,
    num_threads: int = 0,  # pylint: disable=redefined-outer-name
    batch: int = 1024,
) -> tp.Tuple[np.ndarray, np.ndarray]:
    """Same contract as ``implicit.cpu.topk.topk`` (ids int32 [Q,k], scores fp32 [Q,k])."""
    lib = _lib()
    items = np.ascontiguousarray(items, dtype=np.float32)
    query = np.ascontiguousarray(query, dtype=np.float32)
    n_q, n_items = query.shape[0], items.shape[0]
    k = min(int(k), n_items)
    ids = np.empty((n_q, k), dtype=np.int32)
    scores = np.empty((n_q, k), dtype=np.float32)
    norms_p = None
    if item_norms is not None:
        item_norms = np.ascontiguousarray(item_norms, dtype=np.float32).reshape(-1)
        norms_p = item_norms.ctypes.data
    indptr = indices = None
    if filter_query_items is not None:
        indptr = np.ascontiguousarray(filter_query_items.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(filter_query_items.indices, dtype=np.int32)
    items_t = items.T
    for start in range(0, n_q, batch):
        stop = min(start + batch, n_q)
        s = np.ascontiguousarray(query[start:stop] @ items_t)
        rc = lib.ref_topk_select(
            s.ctypes.data, stop - start, n_items, norms_p,
            None if indptr is None else indptr[start:].ctypes.data,
            None if indices is None else indices.ctypes.data,
            k, ids[start:stop].ctypes.data, scores[start:stop].ctypes.data, num_threads,
        )
        if rc != 0:
            raise ValueError("ref_topk_select: bad arguments")
    return ids, scores
