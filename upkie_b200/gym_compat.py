# SPDX-License-Identifier: Apache-2.0
"""Gymnasium when it is installed, a minimal stand-in otherwise.

The reference's envs are ``gymnasium.Env`` subclasses (``upkie/envs/upkie_env.py:19``)
with ``gymnasium.spaces`` (``upkie/envs/upkie_servos.py:173-253``). Gymnasium is
an optional dependency of this package: when it imports, the vector env derives
from ``gymnasium.vector.VectorEnv`` and uses its spaces; otherwise the small
classes below provide the same attributes (``low``, ``high``, ``shape``,
``dtype``, ``spaces``, ``sample``, ``contains``) so that the host logic and the
tests run without it.
"""

from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the environment
    import gymnasium as _gym
    from gymnasium import spaces
    from gymnasium.vector import VectorEnv
    from gymnasium.vector.utils import batch_space

    HAVE_GYMNASIUM = True
    Env = _gym.Env
except ImportError:
    HAVE_GYMNASIUM = False

    class _Space:
        def __init__(self, shape=None, dtype=None, seed=None):
            self._shape = None if shape is None else tuple(shape)
            self.dtype = None if dtype is None else np.dtype(dtype)
            self._np_random = np.random.default_rng(seed)

        @property
        def shape(self):
            return self._shape

        def seed(self, seed=None):
            self._np_random = np.random.default_rng(seed)

        def __contains__(self, x):
            return self.contains(x)

    class Box(_Space):
        def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
            dtype = np.dtype(dtype)
            if shape is None:
                shape = np.shape(low) if np.ndim(low) > 0 else np.shape(high)
            shape = tuple(shape)
            self.low = np.broadcast_to(np.asarray(low, dtype=dtype), shape).copy()
            self.high = np.broadcast_to(np.asarray(high, dtype=dtype), shape).copy()
            super().__init__(shape, dtype, seed)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1e6)
            hi = np.where(np.isfinite(self.high), self.high, 1e6)
            return self._np_random.uniform(lo, hi, size=self.shape).astype(self.dtype)

        def contains(self, x):
            x = np.asarray(x)
            return bool(
                x.shape == self.shape
                and np.can_cast(x.dtype, self.dtype, casting="same_kind")
                and np.all(x >= self.low)
                and np.all(x <= self.high)
            )

        def __repr__(self):
            return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

        def __eq__(self, other):
            return (
                isinstance(other, Box)
                and self.shape == other.shape
                and self.dtype == other.dtype
                and np.array_equal(self.low, other.low)
                and np.array_equal(self.high, other.high)
            )

    class Dict(_Space):
        def __init__(self, spaces=None, seed=None):
            self.spaces = OrderedDict(spaces or {})
            super().__init__(None, None, seed)

        def sample(self):
            return {k: s.sample() for k, s in self.spaces.items()}

        def contains(self, x):
            return (
                isinstance(x, dict)
                and x.keys() == self.spaces.keys()
                and all(self.spaces[k].contains(x[k]) for k in self.spaces)
            )

        def __getitem__(self, key):
            return self.spaces[key]

        def __iter__(self):
            return iter(self.spaces)

        def __len__(self):
            return len(self.spaces)

        def keys(self):
            return self.spaces.keys()

        def items(self):
            return self.spaces.items()

        def __eq__(self, other):
            return isinstance(other, Dict) and self.spaces == other.spaces

        def __repr__(self):
            return "Dict(" + ", ".join(f"{k!r}: {s}" for k, s in self.spaces.items()) + ")"

    class _SpacesModule:
        pass

    spaces = _SpacesModule()
    spaces.Box = Box
    spaces.Dict = Dict
    spaces.Space = _Space

    def batch_space(space, n):
        """Batched version of ``space`` (gymnasium.vector.utils.batch_space)."""
        if isinstance(space, Box):
            low = np.broadcast_to(space.low, (n,) + space.shape).copy()
            high = np.broadcast_to(space.high, (n,) + space.shape).copy()
            return Box(low, high, shape=(n,) + space.shape, dtype=space.dtype)
        if isinstance(space, Dict):
            return Dict({k: batch_space(s, n) for k, s in space.spaces.items()})
        raise TypeError(f"cannot batch {space!r}")

    class Env:
        """Stand-in for ``gymnasium.Env`` (``reset(seed=...)`` seeds ``np_random``)."""

        metadata = {}
        action_space = None
        observation_space = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.default_rng()
            return self._np_random

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

        @property
        def unwrapped(self):
            return self

        def close(self):
            pass

    class VectorEnv:
        """Stand-in for ``gymnasium.vector.VectorEnv``."""

        metadata = {}
        num_envs = 0
        single_action_space = None
        single_observation_space = None
        action_space = None
        observation_space = None
        _np_random = None

        @property
        def np_random(self):
            if self._np_random is None:
                self._np_random = np.random.default_rng()
            return self._np_random

        def reset(self, *, seed=None, options=None):
            if seed is not None:
                self._np_random = np.random.default_rng(seed)

        @property
        def unwrapped(self):
            return self

        def close(self, **kwargs):
            pass
