"""nig_rollout_host (what bench.py's e2e times) over 1 / 2 / 4 / 8 env slices: 65,536 reactor envs x 1,000 steps, K = 64,
pinned host buffers in and out -> env-steps/s per slice count (NIG_HOST_SLICES)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np
import neorl_industrial as ni

n, T = 65536, 1000
env = ni.make("ChemicalReactor-v0", num_envs=n, seed=0, copy=False)
init = env.native.reset_host().copy()
pinned = env.native.pinned("sweep_init", init.shape, np.float32)
pinned[...] = init
for slices in (1, 2, 4, 8, 1, 4):
    os.environ["NIG_HOST_SLICES"] = str(slices)
    for _ in range(5):
        env.rollout(T, "random", steps_per_launch=64, init_states=pinned)
    t0 = time.perf_counter()
    reps = 100
    for _ in range(reps):
        out = env.rollout(T, "random", steps_per_launch=64, init_states=pinned)
    dt = (time.perf_counter() - t0) / reps
    print(f"slices={slices}: {n * T / dt:.4g} env-steps/s ({dt * 1e3:.3f} ms per call) checksum {float(out['reward_sum'].sum()):.6e}", flush=True)
env.close()
