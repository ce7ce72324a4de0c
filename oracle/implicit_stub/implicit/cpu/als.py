import numpy as np


class AlternatingLeastSquares:
    """Attribute carrier only (no solver): enough for ImplicitALSWrapperModel to wrap a pre-fitted model
    (pattern: tests/models/test_implicit_als.py:193-197)."""

    def __init__(self, factors=100, regularization=0.01, alpha=1.0, dtype=np.float32, use_native=True, use_cg=True,
                 iterations=15, calculate_training_loss=False, num_threads=0, random_state=None):
        self.factors = factors
        self.regularization = regularization
        self.alpha = alpha
        self.dtype = np.dtype(dtype)
        self.use_native = use_native
        self.use_cg = use_cg
        self.iterations = iterations
        self.calculate_training_loss = calculate_training_loss
        self.num_threads = num_threads
        self.random_state = random_state
        self.user_factors = None
        self.item_factors = None

    def fit(self, *args, **kwargs):
        raise NotImplementedError("implicit stub: ALS solver is out of scope")
