"""Observation / action spaces.

The reference uses ``gym.spaces.Box`` (gym 0.21, rendezvous_env.py:133-144).  When
gymnasium or gym is importable the real class is used so Stable-Baselines3 accepts
the spaces; otherwise a minimal stand-in with the same ``contains``/``sample``
semantics is provided (neither package is part of this image).
"""
from __future__ import annotations

import numpy as np


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng(seed)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return bool(np.can_cast(x.dtype, self.dtype) and x.shape == self.shape
                    and np.all(x >= self.low) and np.all(x <= self.high))

    def sample(self):
        return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"

    def __eq__(self, other):
        return (hasattr(other, "low") and hasattr(other, "shape") and self.shape == tuple(other.shape)
                and np.allclose(self.low, other.low) and np.allclose(self.high, other.high))


def _find_box():
    for mod in ("gymnasium", "gym"):
        try:
            return __import__(mod + ".spaces", fromlist=["Box"]).Box
        except Exception:
            continue
    return _Box


Box = _find_box()


def observation_space():
    return Box(low=-1, high=1, shape=(17,), dtype=np.float32)


def action_space():
    return Box(low=-1, high=1, shape=(6,), dtype=np.float32)
