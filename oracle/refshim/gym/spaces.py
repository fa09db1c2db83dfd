"""gym.spaces stand-in: Discrete and Box only (test infrastructure)."""
import random as _random

import numpy as _np

# The oracle harness replaces this hook so Discrete.sample() is drawn from the
# replayed counter-based stream instead of the process RNG.
_sample_hook = None


class Discrete:
    def __init__(self, n):
        self.n = int(n)

    def sample(self):
        if _sample_hook is not None:
            return int(_sample_hook(self.n))
        return _random.randrange(self.n)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class Box:
    def __init__(self, low, high, shape=None, dtype=_np.float32):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

    def sample(self):
        return _np.random.uniform(self.low, self.high, size=self.shape).astype(self.dtype)
