"""matplotlib stand-in (test infrastructure): import-only."""


def use(*a, **k):
    pass
