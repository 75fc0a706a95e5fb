"""Where the host-buffer rollout call (nig_rollout_host) spends its time: NIG_HOST_GRAPH / NIG_HOST_SLICES x with / without the copies."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N

def run(graph, slices, with_init, want_obs, n=65536, T=1000, K=64, reps=30):
    os.environ["NIG_HOST_GRAPH"] = str(graph)
    os.environ["NIG_HOST_SLICES"] = str(slices)
    env = ni.NativeEnv(N.ENV_CHEMICAL_REACTOR, n, device=0, seed=0)
    init = env.pinned("init_states", (n, 12), np.float32)
    init[:] = env.reset_host()
    kw = dict(steps_per_launch=K, init_states=init if with_init else None, want_obs=want_obs)
    for _ in range(5): env.rollout_host(T, N.POLICY_UNIFORM, **kw)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); env.rollout_host(T, N.POLICY_UNIFORM, **kw); ts.append(time.perf_counter() - t0)
    lc = env.launch_count
    env.close()
    return np.median(ts) * 1e3, min(ts) * 1e3, lc

if __name__ == "__main__":
    for graph in (0, 1):
        for slices in (4, 8):
            for with_init, want_obs in ((True, True), (True, False), (False, True), (False, False)):
                med, best, lc = run(graph, slices, with_init, want_obs)
                print(f"graph={graph} slices={slices} init_h2d={with_init} obs_d2h={want_obs}: median {med:.3f} ms  best {best:.3f} ms  "
                      f"-> {65536 * 1000 / (med * 1e-3):.4g} env-steps/s  (launches {lc})", flush=True)
