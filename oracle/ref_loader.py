"""Import the UNMODIFIED reference env modules (from /root/reference in the build container, or from the git-ignored
copy oracle/_ref that oracle/make_ref.py makes of the hot-path files so that the reference's own step loop can be timed
on the GPU box's host cores).

Test / measurement infrastructure only (golden-vector generation, oracle pinning, bench.py's CPU baseline legs). The reference's package
``__init__`` drags in jax/flax/optax/gymnasium, none of which are installed here, so we:

  1. register a ~25-line stub ``gymnasium`` (``Env`` + ``spaces.Box``) and a stub ``jax``/``jax.numpy``
     whose only used attribute is ``ndarray`` (``isinstance`` check at environments/base.py:163);
  2. pre-register an empty ``neorl_industrial`` package object pointing at the reference tree so the
     real ``__init__`` (agents, quality gates, ...) never executes;
  3. import ``neorl_industrial.environments`` -- the arithmetic then runs through the real numpy.

Nothing here is imported by the product package. On the GPU box /root/reference does not exist: the parity tests use the
committed .npz fixtures, and only bench.py's CPU baseline legs load the reference, from oracle/_ref.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_COPY_SRC = os.path.join(_HERE, "_ref")          # oracle/make_ref.py output (git-ignored; travels with gpurun)
REFERENCE_SRC = os.environ.get("NIG_REFERENCE_SRC", "/root/reference/src")
_REF_PKG = "neorl_industrial"


def reference_available(src: str = None) -> bool:
    return os.path.isdir(os.path.join(src or REFERENCE_SRC, _REF_PKG, "environments"))


class _Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape) if shape is not None else np.shape(low)
        self.low = np.full(self.shape, low, dtype=self.dtype)
        self.high = np.full(self.shape, high, dtype=self.dtype)
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def sample(self):
        # gymnasium.spaces.Box.sample for a bounded box: uniform in fp64, cast to the box dtype
        return self._rng.uniform(self.low, self.high, self.shape).astype(self.dtype)


class _Env:
    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)


def _install_stubs():
    if "gymnasium" not in sys.modules:
        gym = types.ModuleType("gymnasium")
        spaces = types.ModuleType("gymnasium.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        gym.Env = _Env
        sys.modules["gymnasium"] = gym
        sys.modules["gymnasium.spaces"] = spaces
    if "jax" not in sys.modules:
        jax = types.ModuleType("jax")
        jnp = types.ModuleType("jax.numpy")

        class _NeverArray:  # isinstance(x, jnp.ndarray) must simply be False for numpy inputs
            pass

        jnp.ndarray = _NeverArray
        jax.numpy = jnp
        sys.modules["jax"] = jax
        sys.modules["jax.numpy"] = jnp


def load_reference_envs(src: str = None):
    """Returns the reference's ``neorl_industrial.environments`` module (real code, stubbed deps) from `src`
    (default: the reference tree; oracle/_ref for the copy).

    The product package is also called ``neorl_industrial`` (drop-in), so the reference copy is
    loaded under a private alias and ``sys.modules`` is restored afterwards.
    """
    src = src or REFERENCE_SRC
    if not reference_available(src):
        raise RuntimeError(f"reference tree not found under {src}")
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == _REF_PKG or k.startswith(_REF_PKG + ".")}
    for k in saved:
        del sys.modules[k]
    try:
        pkg = types.ModuleType(_REF_PKG)
        pkg.__path__ = [os.path.join(src, _REF_PKG)]
        sys.modules[_REF_PKG] = pkg
        envs = importlib.import_module(_REF_PKG + ".environments")
        types_mod = importlib.import_module(_REF_PKG + ".core.types")
        envs._ref_types = types_mod
        return envs
    finally:
        for k in [k for k in sys.modules if k == _REF_PKG or k.startswith(_REF_PKG + ".")]:
            del sys.modules[k]
        sys.modules.update(saved)


class NoiseQueue:
    """Teacher-forcing interposer for the reference's global ``np.random.normal/uniform/random``.

    Draw order in the reference: ChemicalReactor step -> [temp_noise, pressure_noise]
    (chemical_reactor.py:149,159); PowerGrid step -> V(8), load(8), flow(7) arrays
    (power_grid.py:136,140,144). Supplied values are *standard-scaled already* (i.e. the value the
    call must return), popped in call order.
    """

    def __init__(self):
        self.queue = []
        self._orig = None

    def push(self, *values):
        self.queue.extend(values)

    def _normal(self, loc=0.0, scale=1.0, size=None):
        v = self.queue.pop(0)
        if size is None:
            return float(v)
        v = np.asarray(v, dtype=np.float64)
        assert v.shape == (size,) or v.shape == tuple(np.atleast_1d(size)), (v.shape, size)
        return v

    def __enter__(self):
        self._orig = np.random.normal
        np.random.normal = self._normal
        return self

    def __exit__(self, *exc):
        np.random.normal = self._orig
        return False
