"""`implicit.cpu.topk.topk` stand-in: forwards to the numpy restatement (oracle/topk_oracle.py)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..")))
from oracle.topk_oracle import implicit_topk  # noqa: E402


def topk(items, query, k, item_norms=None, filter_query_items=None, filter_items=None, num_threads=0):
    return implicit_topk(items, query, k, item_norms, filter_query_items, filter_items, num_threads)
