import numpy as np


class ParameterWarning(Warning):
    pass


def check_random_state(random_state):
    if isinstance(random_state, np.random.RandomState):
        return random_state
    return np.random.RandomState(random_state)
