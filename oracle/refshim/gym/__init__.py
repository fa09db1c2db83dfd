"""Stand-in for the `gym` package (TEST INFRASTRUCTURE ONLY).

The reference imports `gym` only for `gym.Env`, `gym.spaces.Discrete/Box` and
`gym.utils.seeding.np_random` (CyberDefenseEnv.py:1-2,18,27-29,262).  This
stub provides exactly that surface so the unmodified reference modules can be
imported in a container that does not ship gym.  It is never imported by the
product package.
"""
from . import spaces  # noqa: F401
from . import utils  # noqa: F401


class Env:
    metadata = {}

    def __init__(self, *a, **k):
        pass

    def reset(self, *a, **k):
        raise NotImplementedError

    def step(self, *a, **k):
        raise NotImplementedError
