"""imageio stand-in (test infrastructure): import-only."""
