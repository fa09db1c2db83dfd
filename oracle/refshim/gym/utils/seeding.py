import numpy as _np


def np_random(seed=None):
    if seed is None:
        seed = 0
    return _np.random.RandomState(int(seed) % (2 ** 32)), seed
