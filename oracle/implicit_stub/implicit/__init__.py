"""Import-only stand-in for the third-party `implicit` package (pm-implicit 0.7.3), TEST INFRASTRUCTURE ONLY.

`implicit` is a hard import of `rectools.models` (rectools/models/__init__.py:42-45) but is neither vendored under
/root/reference nor installable here (no network).  With this directory on PYTHONPATH the UNMODIFIED reference imports
and its ranking path runs; the only function carrying arithmetic is `implicit.cpu.topk.topk`, which forwards to the
numpy restatement in `oracle/topk_oracle.py`.  Used by `oracle/make_golden.py` to generate `tests/golden/`.
"""
from . import cpu, gpu  # noqa: F401

__version__ = "0.7.3+stub"
