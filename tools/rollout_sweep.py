"""Fused rollout kernel sweep: population x K x CTA size -> env-steps/s (uniform in-kernel policy)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N

def run(kind, n, block, K, horizon=1024, reps=5, policy=N.POLICY_UNIFORM):
    os.environ["NIG_ROLLOUT_BLOCK"] = str(block)
    env = ni.NativeEnv(kind, n, device=0, seed=0)
    env.reset_device()
    def one_pass():
        done = 0
        while done < horizon:
            k = min(K, horizon - done)
            env.rollout_device(k, policy)
            done += k
    for _ in range(2): one_pass()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); one_pass(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    st = env.stats_dict()
    env.close()
    return n * horizon / (ms * 1e-3), ms, st["steps"], st["return_sum"]

if __name__ == "__main__":
    kinds = {"reactor": 0, "grid": 1, "robot": 2}
    which = sys.argv[1:] or ["reactor"]
    for name in which:
        kind = kinds[name]
        for n in (65536, 1 << 18, 1 << 20):
            for K in (64, 256):
                for block in (128,):
                    rate, ms, steps, ret = run(kind, n, block, K)
                    print(f"{name:8s} n={n:8d} K={K:4d} block={block:3d}: {rate:.4g} env-steps/s ({ms:.3f} ms / 1024 steps) "
                          f"steps={steps} return_sum={ret:.9e}", flush=True)
