"""The env-path sections of the reference's timing harness (performance_benchmark.py:25-50, :81-133) on the B200 path.

Same measurements, same result keys -- env construction, ``get_dataset('medium')`` and the 1,000-step
``action_space.sample()`` / ``step`` / reset-on-done loop through the single-env gym API -- plus the batched form of
the step loop (``num_envs > 1``: every env runs the loop inside the fused rollout kernel). The agent / training
sections of the harness are outside the step path and not reproduced.

    python -m neorl_industrial.benchmarks.performance [--num-envs 65536] [--steps 1000]
"""
from __future__ import annotations

import argparse
import json
import os
import resource
import time
from typing import Any, Dict

from ..utils import make


def _rss_mb() -> float:
    try:
        with open(f"/proc/{os.getpid()}/statm") as f:
            return int(f.read().split()[1]) * resource.getpagesize() / 2 ** 20
    except OSError:
        return resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024.0


def benchmark_environment_creation(n_envs: int = 10, env_id: str = "ChemicalReactor-v0") -> Dict[str, float]:
    """performance_benchmark.py:25-50 -- keys avg_creation_time, memory_per_env, total_time."""
    mem0, t0 = _rss_mb(), time.time()
    envs = [make(env_id) for _ in range(n_envs)]
    total = time.time() - t0
    mem1 = _rss_mb()
    for e in envs:
        e.close()
    return {"avg_creation_time": total / n_envs, "memory_per_env": (mem1 - mem0) / n_envs, "total_time": total}


def benchmark_dataset_loading(quality: str = "medium", env_id: str = "ChemicalReactor-v0") -> Dict[str, float]:
    """performance_benchmark.py:81-103 -- keys dataset_size, load_time, samples_per_sec."""
    env = make(env_id)
    import torch
    torch.zeros(1, device=env.native.torch_device())      # torch's one-time CUDA context set-up is not dataset generation
    t0 = time.time()
    dataset = env.get_dataset(quality=quality)
    load_time = time.time() - t0
    n = len(dataset["observations"])
    env.close()
    return {"dataset_size": n, "load_time": load_time, "samples_per_sec": n / load_time}


def benchmark_environment_steps(n_steps: int = 1000, num_envs: int = 1, env_id: str = "ChemicalReactor-v0",
                                steps_per_launch: int = 64) -> Dict[str, float]:
    """performance_benchmark.py:106-133 -- keys total_time, steps_per_sec. ``num_envs == 1`` runs the reference's loop
    verbatim through the gym API (one kernel launch and one PCIe round trip per step); ``num_envs > 1`` runs the same
    loop for every env at once with ``env.rollout`` (host arrays in and out)."""
    if num_envs == 1:
        env = make(env_id)
        env.reset()
        t0 = time.time()
        for _ in range(n_steps):
            _, _, terminated, truncated, _ = env.step(env.action_space.sample())
            if terminated or truncated:
                env.reset()
        total = time.time() - t0
        env.close()
        return {"total_time": total, "steps_per_sec": n_steps / total}
    env = make(env_id, num_envs=num_envs, copy=False)
    env.rollout(steps_per_launch, "random", steps_per_launch=steps_per_launch, reset=True)      # warm-up
    t0 = time.time()
    res = env.rollout(n_steps, "random", steps_per_launch=steps_per_launch, reset=True)
    total = time.time() - t0
    env.close()
    return {"total_time": total, "steps_per_sec": num_envs * n_steps / total, "num_envs": num_envs,
            "episodes": int(res["episodes"].sum()), "violations": int(res["violations"].sum())}


def run_all(num_envs: int = 65536, n_steps: int = 1000) -> Dict[str, Any]:
    return {
        "environment_creation": benchmark_environment_creation(),
        "dataset_loading": benchmark_dataset_loading(),
        "environment_steps": benchmark_environment_steps(n_steps),
        "environment_steps_batched": benchmark_environment_steps(n_steps, num_envs=num_envs),
    }


def main() -> None:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--num-envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=1000)
    args = ap.parse_args()
    print(json.dumps(run_all(args.num_envs, args.steps), indent=2))


if __name__ == "__main__":
    main()
