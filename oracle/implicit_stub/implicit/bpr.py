from .cpu.bpr import BayesianPersonalizedRanking as _CPU


def BayesianPersonalizedRanking(*args, use_gpu=False, **kwargs):  # noqa: N802
    return _CPU(*args, **kwargs)
