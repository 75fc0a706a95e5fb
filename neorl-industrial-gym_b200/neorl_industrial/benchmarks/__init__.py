"""Baseline controllers of the reference's benchmark suite (benchmarks/baseline_agents.py), usable on the host
(``agent.act(obs)``, the reference API) and inside the fused rollout kernel (``env.rollout(n, agent)``)."""
from .baseline_agents import (BaselineAgent, BaselineAgentFactory, ConstantAgent, MPC_Agent, PIDControllerAgent,
                              RandomAgent)

__all__ = ["BaselineAgent", "RandomAgent", "PIDControllerAgent", "MPC_Agent", "ConstantAgent", "BaselineAgentFactory"]
