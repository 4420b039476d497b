"""`observation_space` / `action_space` of the reference env (src/rl/envs/ackermann_env.py:95-108).

The reference builds ``gymnasium.spaces.Box`` objects.  gymnasium is used when it is importable; otherwise a minimal Box with the
same attributes (low, high, shape, dtype, sample, contains) stands in, so the env classes always expose both spaces.
"""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent from the build image
    from gymnasium import Env as GymEnv
    from gymnasium.spaces import Box
    HAVE_GYMNASIUM = True
except Exception:
    HAVE_GYMNASIUM = False

    class GymEnv:  # noqa: D401 - stand-in base class
        """Stand-in for gymnasium.Env (same reset/step contract, no registry)."""
        metadata = {"render_modes": []}

    class Box:
        """Stand-in for gymnasium.spaces.Box: a (possibly unbounded) box in R^n."""

        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            self.dtype = np.dtype(dtype)
            shape = tuple(shape) if shape is not None else np.shape(low)
            self.shape = shape
            self.low = np.full(shape, low, dtype=self.dtype) if np.isscalar(low) else np.asarray(low, self.dtype).reshape(shape)
            self.high = np.full(shape, high, dtype=self.dtype) if np.isscalar(high) else np.asarray(high, self.dtype).reshape(shape)
            self._rng = np.random.default_rng(seed)

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return self._rng.uniform(lo, hi).astype(self.dtype)

        def contains(self, x) -> bool:
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

        def __contains__(self, x):
            return self.contains(x)

        def __eq__(self, other):
            return (isinstance(other, Box) and self.shape == other.shape and self.dtype == other.dtype
                    and np.array_equal(self.low, other.low) and np.array_equal(self.high, other.high))

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def observation_space(obs_dim: int = 79) -> "Box":
    """ackermann_env.py:95-100: Box(-inf, inf, (79,), float32) (72 lidar + odom 3 + goal 4)."""
    return Box(low=-np.inf, high=np.inf, shape=(obs_dim,), dtype=np.float32)


def action_space() -> "Box":
    """ackermann_env.py:103-108: Box(-1, 1, (2,), float32) (linear_x, angular_z)."""
    return Box(low=-1.0, high=1.0, shape=(2,), dtype=np.float32)
