import numpy as np


class BayesianPersonalizedRanking:
    def __init__(self, factors=100, learning_rate=0.01, regularization=0.01, dtype=np.float32, iterations=100,
                 verify_negative_samples=True, num_threads=0, random_state=None):
        self.factors = factors
        self.learning_rate = learning_rate
        self.regularization = regularization
        self.dtype = np.dtype(dtype)
        self.iterations = iterations
        self.verify_negative_samples = verify_negative_samples
        self.num_threads = num_threads
        self.random_state = random_state
        self.user_factors = None
        self.item_factors = None

    def fit(self, *args, **kwargs):
        raise NotImplementedError("implicit stub: BPR solver is out of scope")
