"""gymnasium spaces when gymnasium is importable, otherwise a minimal bundled stand-in with the same
attributes the PCGRL code paths read (n, nvec, low, high, shape, dtype, spaces, sample)."""
from __future__ import annotations

import numpy as np

try:  # pragma: no cover - gymnasium is absent in the build image
    from gymnasium import Env as GymEnv, Wrapper as GymWrapper
    from gymnasium.spaces import Box, Dict, Discrete, MultiDiscrete
    HAVE_GYMNASIUM = True
except Exception:  # noqa: BLE001
    HAVE_GYMNASIUM = False

    class Space:
        def __init__(self, shape=None, dtype=None):
            self.shape = None if shape is None else tuple(int(s) for s in shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._rng = np.random.default_rng()

        def seed(self, seed=None):
            self._rng = np.random.default_rng(seed)
            return [seed]

    class Discrete(Space):
        def __init__(self, n):
            super().__init__((), np.int64)
            self.n = int(n)

        def sample(self):
            return int(self._rng.integers(self.n))

        def contains(self, x):
            return 0 <= int(x) < self.n

        def __repr__(self):
            return f"Discrete({self.n})"

    class MultiDiscrete(Space):
        def __init__(self, nvec):
            self.nvec = np.asarray(nvec, dtype=np.int64)
            super().__init__(self.nvec.shape, np.int64)

        def sample(self):
            return (self._rng.random(self.nvec.shape) * self.nvec).astype(np.int64)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.nvec.shape and bool(np.all((x >= 0) & (x < self.nvec)))

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is None:
                shape = np.broadcast(np.asarray(low), np.asarray(high)).shape
            super().__init__(shape, dtype)
            self.low = np.broadcast_to(np.asarray(low), self.shape).astype(self.dtype)
            self.high = np.broadcast_to(np.asarray(high), self.shape).astype(self.dtype)

        def sample(self):
            u = self._rng.random(self.shape)
            return (self.low + u * (self.high.astype(np.float64) - self.low)).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return x.shape == self.shape and bool(np.all((x >= self.low) & (x <= self.high)))

    class Dict(Space):
        def __init__(self, spaces=None):
            super().__init__(None, None)
            self.spaces = dict(spaces or {})

        def __getitem__(self, k):
            return self.spaces[k]

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

    class GymEnv:
        metadata: dict = {}

        @property
        def unwrapped(self):
            return self

    class GymWrapper(GymEnv):
        def __init__(self, env):
            self.env = env

        def __getattr__(self, name):
            if name.startswith("_"):
                raise AttributeError(name)
            return getattr(self.env, name)

        @property
        def unwrapped(self):
            return self.env.unwrapped

        def reset(self, *, seed=None, options=None):
            return self.env.reset(seed=seed, options=options)

        def step(self, action):
            return self.env.step(action)
