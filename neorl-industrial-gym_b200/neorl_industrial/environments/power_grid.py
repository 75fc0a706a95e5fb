"""PowerGrid-v0 (reference environments/power_grid.py): 32-d state, 8-d action. Physics: csrc/nig_envs.cuh (struct Grid)."""
from __future__ import annotations

from typing import Dict

import numpy as np

from .. import _native as N
from ..core.types import SafetyConstraint
from ..datasets import episodes_for_transitions, generate_dataset
from .base import IndustrialEnv


def frequency_constraint(state, action) -> bool:
    """power_grid.py:10-14"""
    return bool(abs(state[0]) < 0.5)


def voltage_constraint(state, action) -> bool:
    """power_grid.py:17-21"""
    v = state[1:9]
    return bool(np.all((v >= 0.95) & (v <= 1.05)))


def generation_constraint(state, action) -> bool:
    """power_grid.py:24-30"""
    new_gen = state[9:17] + action
    return bool(np.all((new_gen >= 0) & (new_gen <= np.ones(8) * 100)))


class PowerGridEnv(IndustrialEnv):
    ENV_ID = "PowerGrid-v0"
    ENV_KIND = N.ENV_POWER_GRID
    DATASET_SAMPLES = {"expert": 100000, "medium": 150000, "mixed": 200000, "random": 80000}   # power_grid.py:197-202

    def __init__(self, **kwargs):
        self.base_load = np.array([50, 60, 45, 55, 40, 65, 35, 50])      # power_grid.py:82-88
        self.load_variation = 0.2
        self.inertia_constant = 5.0
        self.damping_factor = 1.0
        self.generation_cost = np.array([25, 30, 28, 35, 32, 27, 40, 33])
        constraints = [   # power_grid.py:53-72
            SafetyConstraint("frequency_stability", frequency_constraint, -50.0, True, _native=("builtin", 0)),
            SafetyConstraint("voltage_limits", voltage_constraint, -30.0, True, _native=("builtin", 1)),
            SafetyConstraint("generation_limits", generation_constraint, -20.0, False, _native=("builtin", 2)),
        ]
        kwargs.setdefault("max_episode_steps", 1000)
        kwargs.setdefault("dt", 0.1)
        super().__init__(state_dim=32, action_dim=8, safety_constraints=constraints, **kwargs)

    @classmethod
    def dataset_policy(cls, quality: str):
        """power_grid.py:197-232; unknown quality -> KeyError like the reference."""
        n_samples = cls.DATASET_SAMPLES[quality]
        pp = N.PolicyParams()
        pp.store_clip, pp.mode = 0.0, 0
        if quality == "expert":
            pp.p_ctrl, pp.uniform_scale = 1.0, 1.0
            for k in range(8):
                pp.gain[k][0], pp.gain[k][1] = -0.5, 0.1
        elif quality == "random":
            pp.p_ctrl, pp.uniform_scale = 0.0, 5.0
        else:
            pp.p_ctrl, pp.uniform_scale = 0.6, 3.0
            for k in range(8):
                pp.gain[k][0] = -0.3
        return n_samples // 1000, 1000, N.POLICY_PCTRL, pp

    def get_dataset(self, quality: str = "mixed", *, n_episodes=None, n_transitions=None, extensions: bool = False) -> Dict[str, np.ndarray]:
        n_ep, n_steps, policy, pp = self.dataset_policy(quality)
        if n_transitions is not None:      # as many whole episodes as it takes to reach n_transitions
            n_episodes = episodes_for_transitions(self.native, int(n_transitions), n_steps, policy, pp)
        return generate_dataset(self, n_episodes or n_ep, n_steps, policy, pp, terminals_include_truncation=False,
                                timeouts_key=False, extensions=extensions)
