"""matplotlib.pyplot stand-in (test infrastructure): plotting is out of scope."""


def _noop(*a, **k):
    return None


def __getattr__(name):
    return _noop
