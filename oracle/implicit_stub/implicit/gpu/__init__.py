HAS_CUDA = False
from . import als, bpr  # noqa: E402,F401


class Matrix:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("implicit stub: no GPU support")


COOMatrix = Matrix


class KnnQuery:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("implicit stub: no GPU support")
