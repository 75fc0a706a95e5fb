"""``neorl_industrial.safety`` -- the SafetyWrapper of the reference README (README.md:126-139).

The reference ships no code for it; semantics are defined from the README contract plus the built-in
constraint machinery (environments/base.py:94-124, :170, :179-183): every extra constraint is evaluated
on the PRE-step state and the clipped action exactly like a built-in one; a violation adds ``penalty``
to the reward and counts as a violation; ``critical=True`` also triggers the critical shutdown.

Constraints come in two forms:
  * declarative ``BoundConstraint`` (lo <= s[i] + coef*a[j] <= hi): evaluated inside the CUDA kernels, usable
    in fused rollouts and at any batch size -- the form for temperature / pressure bounds;
  * arbitrary callables ``fn(state, action) -> bool``: evaluated on the host copy of the pre-step state, the
    kernel receives a per-env violation bit. Works for the step API; rejected by fused rollouts.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Union

import numpy as np

from .core.types import SafetyConstraint


def BoundConstraint(name: str, state_index: int, low: float = -np.inf, high: float = np.inf, *, penalty: float = -100.0,
                    critical: bool = False, action_index: int = -1, action_coef: float = 0.0,
                    description: str = "") -> SafetyConstraint:
    """``low <= state[state_index] + action_coef * action[action_index] <= high`` (fp32, unfused), in-kernel."""
    lo, hi, coef = np.float32(low), np.float32(high), np.float32(action_coef)

    def check(state, action) -> bool:
        v = np.float32(state[state_index])
        if action_index >= 0:
            v = np.float32(v + np.float32(coef * np.float32(action[action_index])))
        return bool(lo <= v <= hi)

    return SafetyConstraint(name, check, float(penalty), bool(critical), description,
                            _native=("bound", int(state_index), int(action_index), float(coef), float(lo), float(hi)))


class SafetyWrapper:
    """``SafetyWrapper(env, constraints=[...], penalty=-100)``: env with extra safety constraints.

    ``constraints`` items may be ``SafetyConstraint`` objects (incl. ``BoundConstraint(...)``) or plain
    callables ``fn(state, action) -> bool`` (given ``penalty``, non-critical). Everything else is delegated
    to the wrapped env, so the wrapper is itself a drop-in env."""

    def __init__(self, env, constraints: Sequence[Union[SafetyConstraint, Callable]] = (), penalty: float = -100.0,
                 critical: bool = False):
        self.env = env
        self.penalty = float(penalty)
        self._added: List[str] = []
        for k, c in enumerate(constraints):
            if not isinstance(c, SafetyConstraint):
                if not callable(c):
                    raise TypeError(f"constraint {k} is neither a SafetyConstraint nor callable")
                c = SafetyConstraint(getattr(c, "__name__", f"wrapper_constraint_{k}"), c, self.penalty, bool(critical),
                                     (c.__doc__ or "").strip())
            env.add_safety_constraint(c)
            self._added.append(c.name)

    def __getattr__(self, name):
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action, **kwargs):
        return self.env.step(action, **kwargs)

    def get_dataset(self, quality: str = "mixed", **kwargs):
        return self.env.get_dataset(quality, **kwargs)

    def unwrap(self):
        for name in self._added:
            self.env.remove_safety_constraint(name)
        self._added = []
        return self.env

    @property
    def unwrapped(self):
        return self.env
