from . import als, bpr, matrix_factorization_base, topk  # noqa: F401
