"""nashpy stand-in (test infrastructure): import-only."""


class Game:  # pragma: no cover - never exercised by the step path
    def __init__(self, *a, **k):
        raise NotImplementedError("nashpy stand-in: Nash solving is outside the hot path")
