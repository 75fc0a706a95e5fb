"""Times the reference's OWN environment step loop (performance_benchmark.py:106-133, verbatim) on this machine's cores.

Measurement infrastructure for bench.py's `cpu_baseline` / `--impl reference` legs; runs the unmodified reference modules
from oracle/_ref (oracle/make_ref.py) through the dependency stubs of oracle/ref_loader.py (gymnasium / jax are not in
the image; the arithmetic runs through the real numpy). Always executed in a process of its own (bench.py spawns it), so
the workers are forked from an interpreter that never touched CUDA.

    python oracle/ref_bench.py --mode single --steps 1000 --repeats 5
    python oracle/ref_bench.py --mode all --procs P --steps 1000 --repeats 3      # P forked processes, one env each

Prints one JSON object. The reference has no batched / jit / vmap path for this loop.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)          # sibling modules (the directory name `oracle` is shadowed by oracle.py when run as a script)

ENV_CLASS = {"ChemicalReactor-v0": "ChemicalReactorEnv", "PowerGrid-v0": "PowerGridEnv", "RobotAssembly-v0": "RobotAssemblyEnv"}


def load_envs():
    import make_ref
    import ref_loader
    if make_ref.available():
        return ref_loader.load_reference_envs(make_ref.DEST), "oracle/_ref"
    if ref_loader.reference_available():
        return ref_loader.load_reference_envs(), ref_loader.REFERENCE_SRC
    raise RuntimeError("neither oracle/_ref nor the reference tree is present")


def step_loop(env, n_steps: int) -> float:
    """performance_benchmark.py:106-133 -- the loop between start_time and end_time, unchanged."""
    obs, _ = env.reset()
    start_time = time.perf_counter()
    for i in range(n_steps):
        action = env.action_space.sample()
        obs, reward, terminated, truncated, info = env.step(action)
        if terminated or truncated:
            obs, _ = env.reset()
    return time.perf_counter() - start_time


def _worker(env_id, n_steps, repeats, n_rounds, seed, barrier, out_q):
    import numpy as np
    envs, _ = load_envs()
    np.random.seed(seed)
    env = getattr(envs, ENV_CLASS[env_id])()
    step_loop(env, min(200, n_steps))                 # warm-up (imports, caches)
    spans = []
    for _ in range(n_rounds):                         # one round = one "bench step": every process runs repeats x n_steps
        barrier.wait()
        t0 = time.perf_counter()
        for _ in range(repeats):
            step_loop(env, n_steps)
        spans.append((t0, time.perf_counter()))
    out_q.put(spans)


def run_all(env_id: str, procs: int, n_steps: int, repeats: int, n_rounds: int = 1):
    """P forked processes, one reference env each; returns the wall time of every round (latest end - earliest start)."""
    ctx = mp.get_context("fork")
    barrier, q = ctx.Barrier(procs), ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(env_id, n_steps, repeats, n_rounds, 1000 + i, barrier, q)) for i in range(procs)]
    for p in ps:
        p.start()
    spans = [q.get() for _ in ps]
    for p in ps:
        p.join()
    # perf_counter is CLOCK_MONOTONIC on Linux: comparable across the forked processes
    return [max(w[r][1] for w in spans) - min(w[r][0] for w in spans) for r in range(n_rounds)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", choices=["single", "all"], default="single")
    ap.add_argument("--env", default="ChemicalReactor-v0")
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--repeats", type=int, default=5)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--rounds", type=int, default=1, help="mode all: timed rounds of repeats x steps per process")
    a = ap.parse_args()
    import numpy as np
    envs, src = load_envs()
    out = {"env": a.env, "source": src, "numpy": np.__version__, "loop": "performance_benchmark.py:106-133"}
    if a.mode == "single":
        np.random.seed(0)
        env = getattr(envs, ENV_CLASS[a.env])()
        step_loop(env, min(200, a.steps))
        rates = sorted(a.steps / step_loop(env, a.steps) for _ in range(max(1, a.repeats)))
        out.update({"cores": 1, "steps_per_sec": rates[len(rates) // 2], "min": rates[0], "max": rates[-1],
                    "repeats": len(rates), "steps": a.steps})
    else:
        procs = a.procs or len(os.sched_getaffinity(0))
        walls = run_all(a.env, procs, a.steps, a.repeats, a.rounds)
        out.update({"cores": procs, "steps_per_sec": procs * a.steps * a.repeats * len(walls) / sum(walls), "round_wall_s": walls,
                    "steps": a.steps, "repeats": a.repeats, "env_steps_per_round": procs * a.steps * a.repeats})
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
