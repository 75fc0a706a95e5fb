"""ChemicalReactor-v0 (reference environments/chemical_reactor.py): 12-d state, 3-d action.

Host-side mirror of the reference class: constants, constraint objects, ``_get_safety_info`` and the
``get_dataset`` front-end. The physics lives in csrc/nig_envs.cuh (struct Reactor).
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .. import _native as N
from ..core.types import SafetyConstraint
from ..datasets import episodes_for_transitions, generate_dataset
from .base import IndustrialEnv


class ChemicalReactorEnv(IndustrialEnv):
    ENV_ID = "ChemicalReactor-v0"
    ENV_KIND = N.ENV_CHEMICAL_REACTOR
    REWARD_IS_FLOAT32 = True

    # get_dataset qualities -> (episodes, max steps, noise level)   chemical_reactor.py:333-347
    DATASET_QUALITIES = {"expert": (100, 400, 0.1), "medium": (200, 350, 0.3), "mixed": (300, 300, 0.5), "random": (500, 200, 1.0)}

    def __init__(self, **kwargs):
        # physical parameters (chemical_reactor.py:72-87)
        self.temp_min, self.temp_max, self.temp_target = 280.0, 350.0, 320.0
        self.pressure_min, self.pressure_max, self.pressure_target = 101325.0, 506625.0, 253312.5
        self.reaction_rate_constant = 0.1
        self.heat_capacity = 4.18e3
        self.reactor_volume = 1000.0
        self.temp_noise_std = 1.0
        self.pressure_noise_std = 5000.0
        constraints = [   # chemical_reactor.py:38-60
            SafetyConstraint("temperature_limit", self._temperature_constraint, -100.0, True,
                             "Reactor temperature must stay below 350K", _native=("builtin", 0)),
            SafetyConstraint("pressure_limit", self._pressure_constraint, -50.0, True,
                             "Reactor pressure must stay below 5 atm", _native=("builtin", 1)),
            SafetyConstraint("level_safety", self._level_constraint, -25.0, False,
                             "Reactor level must stay between 20-90%", _native=("builtin", 2)),
        ]
        kwargs.setdefault("max_episode_steps", 500)        # chemical_reactor.py:66 (overridable here, like the other envs)
        kwargs.setdefault("dt", 0.1)
        super().__init__(state_dim=12, action_dim=3, safety_constraints=constraints, **kwargs)

    # host-callable forms of the built-in checks (API parity: tests call constraint.check_fn(obs, action))
    def _temperature_constraint(self, state, action) -> bool:
        temp = state[0] if len(state.shape) == 1 else state[0, 0]
        return bool(temp <= np.float32(self.temp_max))

    def _pressure_constraint(self, state, action) -> bool:
        pressure = state[1] if len(state.shape) == 1 else state[1, 0]
        return bool(pressure <= np.float32(self.pressure_max))

    def _level_constraint(self, state, action) -> bool:
        level = state[10] if len(state.shape) == 1 else state[10, 0]
        return bool(20 <= level <= 90)

    def _get_safety_info(self, state) -> Dict:
        """chemical_reactor.py:307-322"""
        return {
            "safety_metrics": {"temperature": state[0], "pressure": state[1], "level": state[10],
                               "emergency_stop": state[8], "alarm_status": state[9]},
            "constraint_values": {"temp_margin": self.temp_max - state[0], "pressure_margin": self.pressure_max - state[1],
                                  "level_in_bounds": bool(20 <= state[10] <= 90)},
        }

    @classmethod
    def dataset_policy(cls, quality: str):
        """(episodes, steps, policy id, PolicyParams) of chemical_reactor.py:333-390; unknown quality == 'random'."""
        n_ep, n_steps, noise = cls.DATASET_QUALITIES.get(quality, cls.DATASET_QUALITIES["random"])
        pp = N.PolicyParams()
        pp.uniform_scale, pp.store_clip, pp.mode = 1.0, 1.0, 0
        if quality == "expert":
            pp.p_ctrl = 1.0
            pp.gain[0][0], pp.gain[1][0], pp.gain[2][1] = -0.5, 0.3, -0.2          # :370-374
            for k in range(3):
                pp.sigma[k] = noise * 0.1
        else:
            pp.p_ctrl = 1.0 - noise                                               # :378
            pp.gain[0][0] = -0.2                                                  # :381-385
            pp.sigma[0], pp.sigma[1], pp.sigma[2] = noise * 0.3, noise * 0.5, noise * 0.3
        return n_ep, n_steps, N.POLICY_PCTRL, pp

    def get_dataset(self, quality: str = "mixed", *, n_episodes=None, n_transitions=None, extensions: bool = False) -> Dict[str, np.ndarray]:
        """chemical_reactor.py:324-420, generated on the device in D4RL layout."""
        n_ep, n_steps, policy, pp = self.dataset_policy(quality)
        if n_transitions is not None:      # as many whole episodes as it takes to reach n_transitions
            n_episodes = episodes_for_transitions(self.native, int(n_transitions), n_steps, policy, pp)
        return generate_dataset(self, n_episodes or n_ep, n_steps, policy, pp, terminals_include_truncation=True,
                                timeouts_key=True, extensions=extensions)
