"""`gymnasium.spaces.Box` when gymnasium is installed, otherwise a minimal stand-in with the same fields."""
from __future__ import annotations

try:  # pragma: no cover - depends on the environment
    from gymnasium.spaces import Box  # type: ignore
except Exception:

    class Box:  # type: ignore
        def __init__(self, low, high, shape=None, dtype=None):
            self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), dtype

        def __repr__(self):
            return f"Box({self.low}, {self.high}, {self.shape}, {self.dtype})"
