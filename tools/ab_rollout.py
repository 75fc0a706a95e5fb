"""Headline-shaped throughput of whichever library NIG_LIB_PATH points to: ChemicalReactor-v0, uniform policy, K = 64.
   python tools/ab_rollout.py [n_envs ...]   (default 65536 1048576)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N

def run(n, horizon=1000, K=64, reps=15, sliced=True):
    env = ni.NativeEnv(N.ENV_CHEMICAL_REACTOR, n, device=0, seed=0)
    env.reset_device()
    def one_pass():
        if sliced:
            env.rollout_steps_device(horizon, K, N.POLICY_UNIFORM)
        else:
            done = 0
            while done < horizon:
                k = min(K, horizon - done); env.rollout_device(k, N.POLICY_UNIFORM); done += k
    for _ in range(3): one_pass()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); one_pass(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    st = env.stats_dict()
    env.close()
    return n * horizon / (np.median(ts) * 1e-3), n * horizon / (min(ts) * 1e-3), st["return_sum"]

if __name__ == "__main__":
    ns = [int(x) for x in sys.argv[1:]] or [65536, 1 << 20]
    tag = os.path.basename(os.environ.get("NIG_LIB_PATH", "libnig_b200.so"))
    for n in ns:
        for sliced in (True, False):
            med, best, ret = run(n, sliced=sliced)
            print(f"{tag:32s} n={n:8d} {'slices ' if sliced else 'single '} median {med:.4g}  best {best:.4g} env-steps/s  return_sum {ret:.6e}", flush=True)
