"""RobotAssembly-v0 (reference environments/robot_assembly.py): 24-d state, 7-d action. Physics: csrc/nig_envs.cuh (struct Robot)."""
from __future__ import annotations

from typing import Dict

import numpy as np

from .. import _native as N
from ..core.types import SafetyConstraint
from ..datasets import episodes_for_transitions, generate_dataset
from .base import IndustrialEnv


def force_constraint(state, action) -> bool:
    """robot_assembly.py:10-15"""
    return bool(np.all(np.abs(state[18:21]) < 50.0))


def collision_constraint(state, action) -> bool:
    """robot_assembly.py:18-25"""
    p = state[0:3]
    return bool(np.all((p >= np.array([-0.5, -0.5, 0.0])) & (p <= np.array([0.5, 0.5, 0.8]))))


def velocity_constraint(state, action) -> bool:
    """robot_assembly.py:28-32"""
    return bool(np.all(np.abs(state[7:14]) < 2.0))


class RobotAssemblyEnv(IndustrialEnv):
    ENV_ID = "RobotAssembly-v0"
    ENV_KIND = N.ENV_ROBOT_ASSEMBLY
    DATASET_SAMPLES = {"expert": 120000, "medium": 180000, "mixed": 250000, "random": 100000}  # robot_assembly.py:248-253

    def __init__(self, **kwargs):
        self.link_lengths = np.array([0.3, 0.3, 0.25, 0.25, 0.15, 0.1, 0.05])   # robot_assembly.py:85-92
        self.joint_limits_low = np.array([-np.pi] * 7)
        self.joint_limits_high = np.array([np.pi] * 7)
        self.target_position = np.array([0.3, 0.0, 0.4])
        self.insertion_depth = 0.05
        self.alignment_tolerance = 0.005
        constraints = [   # robot_assembly.py:56-75
            SafetyConstraint("force_limits", force_constraint, -100.0, True, _native=("builtin", 0)),
            SafetyConstraint("collision_avoidance", collision_constraint, -200.0, True, _native=("builtin", 1)),
            SafetyConstraint("velocity_limits", velocity_constraint, -50.0, False, _native=("builtin", 2)),
        ]
        kwargs.setdefault("max_episode_steps", 1000)
        kwargs.setdefault("dt", 0.1)
        super().__init__(state_dim=24, action_dim=7, safety_constraints=constraints, **kwargs)

    @classmethod
    def dataset_policy(cls, quality: str):
        """robot_assembly.py:248-291; unknown quality -> KeyError like the reference."""
        n_samples = cls.DATASET_SAMPLES[quality]
        pp = N.PolicyParams()
        pp.store_clip = 2.0
        if quality == "expert":
            pp.p_ctrl, pp.uniform_scale, pp.mode = 1.0, 1.0, 0
            pp.gain[0][0], pp.gain[3][0] = 2.0, -0.1
        elif quality == "random":
            pp.p_ctrl, pp.uniform_scale, pp.mode = 0.0, 1.0, 1
        else:
            pp.p_ctrl, pp.uniform_scale, pp.mode = 0.7, 0.8, 1
            pp.gain[0][0] = 1.0
            pp.sigma[3] = 0.5
        return n_samples // 1000, 1000, N.POLICY_PCTRL, pp

    def get_dataset(self, quality: str = "mixed", *, n_episodes=None, n_transitions=None, extensions: bool = False) -> Dict[str, np.ndarray]:
        n_ep, n_steps, policy, pp = self.dataset_policy(quality)
        if n_transitions is not None:      # as many whole episodes as it takes to reach n_transitions
            n_episodes = episodes_for_transitions(self.native, int(n_transitions), n_steps, policy, pp)
        return generate_dataset(self, n_episodes or n_ep, n_steps, policy, pp, terminals_include_truncation=False,
                                timeouts_key=False, extensions=extensions)
