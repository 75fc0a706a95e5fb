"""Minimal ``Box`` space (gymnasium is not a dependency of the step path; the reference only uses
``Box.low/high/shape/dtype/sample`` -- environments/base.py:60-72, :167)."""
from __future__ import annotations

import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32, seed=None):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else tuple(np.shape(low))
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng(seed)

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        bounded = np.isfinite(self.low) & np.isfinite(self.high)
        out = self._rng.standard_normal(self.shape)
        lo = np.where(bounded, self.low, 0.0)
        hi = np.where(bounded, self.high, 1.0)
        out = np.where(bounded, self._rng.uniform(lo, hi, self.shape), out)
        return out.astype(self.dtype)

    def contains(self, x) -> bool:
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    __contains__ = contains

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"
