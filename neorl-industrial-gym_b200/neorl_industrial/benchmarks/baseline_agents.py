"""Baseline agents (reference: src/neorl_industrial/benchmarks/baseline_agents.py:11-135).

Same classes, constructor arguments and ``act(state)`` results as upstream. In addition every agent describes itself
to the CUDA path (``device_policy()`` -> ``(NIG_POLICY_BASELINE, PolicyParams)``) so that
``env.rollout(n_steps, agent)`` / ``rollouts.evaluate_policy_device`` evaluate it INSIDE the fused K-step kernel for
all envs at once (fp64 controller arithmetic like numpy's, action rounded to fp32; the PID's integral and previous
error live on the device and persist across episodes and calls like the agent object's do --
``env.native.reset_policy_state()`` is the equivalent of constructing a new agent).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional

import numpy as np

from .. import _native as N


class BaselineAgent(ABC):
    """baseline_agents.py:11-25"""

    def __init__(self, state_dim: int, action_dim: int):
        self.state_dim = state_dim
        self.action_dim = action_dim

    @abstractmethod
    def act(self, state: np.ndarray) -> np.ndarray:
        """Select action given state."""

    def predict(self, state: np.ndarray, deterministic: bool = True) -> np.ndarray:
        """``agent.predict`` as evaluate_with_safety calls it (utils.py:96); batched over a leading env axis."""
        state = np.asarray(state)
        if state.ndim == 1:
            return self.act(state)
        return np.stack([self.act(s) for s in state])

    def train(self, dataset: Dict[str, np.ndarray], **kwargs) -> Dict[str, Any]:
        return {"training_complete": True}

    def _params(self, kind: int) -> N.PolicyParams:
        pp = N.PolicyParams()
        pp.baseline.kind = kind
        return pp

    @abstractmethod
    def device_policy(self):
        """(policy id, PolicyParams) of this controller for the fused rollout kernel."""


class RandomAgent(BaselineAgent):
    """baseline_agents.py:28-43"""

    def __init__(self, state_dim: int, action_dim: int, action_low: float = -1.0, action_high: float = 1.0):
        super().__init__(state_dim, action_dim)
        self.action_low = action_low
        self.action_high = action_high

    def act(self, state: np.ndarray) -> np.ndarray:
        return np.random.uniform(self.action_low, self.action_high, size=(self.action_dim,))

    def device_policy(self):
        pp = self._params(N.BASELINE_RANDOM)
        pp.baseline.setpoint[0], pp.baseline.setpoint[1] = float(self.action_low), float(self.action_high)
        return N.POLICY_BASELINE, pp


class PIDControllerAgent(BaselineAgent):
    """baseline_agents.py:46-81"""

    def __init__(self, state_dim: int, action_dim: int, kp: float = 1.0, ki: float = 0.1, kd: float = 0.01,
                 setpoint: Optional[np.ndarray] = None):
        super().__init__(state_dim, action_dim)
        self.kp, self.ki, self.kd = kp, ki, kd
        self.setpoint = setpoint if setpoint is not None else np.zeros(action_dim)
        self.previous_error = np.zeros(action_dim)
        self.integral = np.zeros(action_dim)

    def act(self, state: np.ndarray) -> np.ndarray:
        pv = state[:self.action_dim] if len(state) >= self.action_dim else state
        error = self.setpoint - pv
        proportional = self.kp * error
        self.integral += error
        integral_term = self.ki * self.integral
        derivative = self.kd * (error - self.previous_error)
        action = np.clip(proportional + integral_term + derivative, -1.0, 1.0)
        self.previous_error = error
        return action

    def device_policy(self):
        if self.action_dim > 8:
            raise ValueError("the device PID supports at most 8 action dimensions")
        pp = self._params(N.BASELINE_PID)
        pp.baseline.kp, pp.baseline.ki, pp.baseline.kd = float(self.kp), float(self.ki), float(self.kd)
        for k, v in enumerate(np.asarray(self.setpoint, np.float64).reshape(-1)[:self.action_dim]):
            pp.baseline.setpoint[k] = float(v)
        return N.POLICY_BASELINE, pp


class MPC_Agent(BaselineAgent):
    """baseline_agents.py:84-100 (the "simplified MPC" heuristic: half the negated state)"""

    def __init__(self, state_dim: int, action_dim: int, horizon: int = 10, cost_weights: Optional[np.ndarray] = None):
        super().__init__(state_dim, action_dim)
        self.horizon = horizon
        self.cost_weights = cost_weights if cost_weights is not None else np.ones(action_dim)

    def act(self, state: np.ndarray) -> np.ndarray:
        state_error = np.zeros_like(state) - state
        action = 0.5 * state_error[:self.action_dim] if len(state) >= self.action_dim else np.zeros(self.action_dim)
        return np.clip(action, -1.0, 1.0)

    def device_policy(self):
        return N.POLICY_BASELINE, self._params(N.BASELINE_MPC)


class ConstantAgent(BaselineAgent):
    """baseline_agents.py:103-114"""

    def __init__(self, state_dim: int, action_dim: int, constant_action: Optional[np.ndarray] = None):
        super().__init__(state_dim, action_dim)
        self.constant_action = constant_action if constant_action is not None else np.zeros(action_dim)

    def act(self, state: np.ndarray) -> np.ndarray:
        return self.constant_action.copy()

    def device_policy(self):
        pp = self._params(N.BASELINE_CONSTANT)
        for k, v in enumerate(np.asarray(self.constant_action, np.float64).reshape(-1)[:self.action_dim]):
            pp.baseline.setpoint[k] = float(v)
        return N.POLICY_BASELINE, pp


class BaselineAgentFactory:
    """baseline_agents.py:117-135"""

    AGENTS = {"random": RandomAgent, "pid": PIDControllerAgent, "mpc": MPC_Agent, "constant": ConstantAgent}

    @classmethod
    def create(cls, agent_type: str, state_dim: int, action_dim: int, **kwargs) -> BaselineAgent:
        if agent_type not in cls.AGENTS:
            raise ValueError(f"Unknown baseline agent type: {agent_type}. Available: {list(cls.AGENTS.keys())}")
        return cls.AGENTS[agent_type](state_dim, action_dim, **kwargs)
