from .cpu.als import AlternatingLeastSquares as _CPU


def AlternatingLeastSquares(*args, use_gpu=False, **kwargs):  # noqa: N802
    return _CPU(*args, **kwargs)
