"""``make`` and ``evaluate_with_safety`` (reference utils.py:12-154)."""
from __future__ import annotations

from typing import Any, Dict

import numpy as np

from .environments import ChemicalReactorEnv, PowerGridEnv, RobotAssemblyEnv

_REGISTRY = {
    "ChemicalReactor-v0": ChemicalReactorEnv,
    "PowerGrid-v0": PowerGridEnv,
    "RobotAssembly-v0": RobotAssemblyEnv,
}
# registered upstream (utils.py:30-31) but not instantiable there (abstract methods missing)
_UPSTREAM_BROKEN = ("AdvancedChemicalReactor-v0", "AdvancedPowerGrid-v0")


def make(env_id: str, **kwargs) -> Any:
    """Create an industrial environment by id (utils.py:12-39).

    Extra keyword arguments of the B200 build: ``num_envs`` (default 1), ``device`` ('cuda', 'cuda:N'),
    ``seed``, ``auto_reset``, ``env_id_offset`` (global id of local env 0 when sharding), ``copy``.
    """
    if env_id in _UPSTREAM_BROKEN:
        raise NotImplementedError(f"'{env_id}' is registered upstream but cannot be instantiated there either "
                                  "(abstract methods missing); it is not part of the B200 step path")
    if env_id not in _REGISTRY:
        available = ", ".join(list(_REGISTRY) + list(_UPSTREAM_BROKEN))
        raise ValueError(f"Unknown environment '{env_id}'. Available: {available}")
    return _REGISTRY[env_id](**kwargs)


def evaluate_with_safety(agent: Any, env: Any, n_episodes: int = 100, record_video: bool = False,
                         render: bool = False) -> Dict[str, Any]:
    """Evaluate an agent with safety metrics (utils.py:42-154). Works with the single-env API exactly like the
    reference loop; for batched envs see ``neorl_industrial.rollouts.evaluate_policy_batched``."""
    if not hasattr(agent, "is_trained") or not agent.is_trained:
        raise RuntimeError("Agent must be trained before evaluation")
    if getattr(env, "batched", False):
        from .rollouts import evaluate_agent_batched
        return evaluate_agent_batched(agent, env, n_episodes)
    episode_returns, episode_lengths, satisfaction = [], [], []
    total_violations = critical_violations = emergency_shutdowns = 0
    for _ in range(n_episodes):
        obs, info = env.reset()
        ep_return, ep_len, done = 0.0, 0, False
        while not done:
            action = agent.predict(obs[None], deterministic=True)[0]
            obs, reward, terminated, truncated, info = env.step(action)
            done = terminated or truncated
            ep_return += reward
            ep_len += 1
            sm = info.get("safety_metrics")
            if sm is not None:
                total_violations += sm.violation_count
                critical_violations += sm.critical_violations
                satisfaction.append(sm.satisfaction_rate)
            if info.get("critical_shutdown", False):
                emergency_shutdowns += 1
        episode_returns.append(ep_return)
        episode_lengths.append(ep_len)
    successes = sum(1 for r in episode_returns if r > 0)
    return {
        "return_mean": np.mean(episode_returns), "return_std": np.std(episode_returns),
        "return_min": np.min(episode_returns), "return_max": np.max(episode_returns),
        "length_mean": np.mean(episode_lengths), "length_std": np.std(episode_lengths),
        "safety_violations": total_violations, "safety_violations_per_episode": total_violations / n_episodes,
        "critical_violations": critical_violations, "emergency_shutdowns": emergency_shutdowns,
        "constraint_satisfaction_rate": np.mean(satisfaction) if satisfaction else 1.0,
        "successful_episodes": successes, "success_rate": successes / n_episodes,
    }
