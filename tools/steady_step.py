"""Single-step kernel in STEADY STATE (episodes ending at the natural rate, in-kernel auto-reset firing) vs fresh envs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
dev = torch.device("cuda", 0)
for n in (1 << 22, 1 << 24):
    env = ni.NativeEnv(0, n, device=0, seed=0); env.reset_device()
    acts = torch.rand((3, env.pitch), device=dev) * 2 - 1
    rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
    def t(reps=20):
        for _ in range(3): env.step_device(acts, reward=rew, flags=fl, viol_mask=vm)
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(); env.step_device(acts, reward=rew, flags=fl, viol_mask=vm); b.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs]))
    fresh = t()
    for _ in range(10): env.rollout_device(64, N.POLICY_UNIFORM)      # 640 steps: past the first wave of episode ends
    torch.cuda.synchronize()
    c0 = env.stats_dict()["episodes"]
    steady = t()
    c1 = env.stats_dict()["episodes"]
    print(f"n={n}: fresh {122*n/fresh/1e6:.0f} GB/s ({fresh*1e3:.1f} us), steady {122*n/steady/1e6:.0f} GB/s ({steady*1e3:.1f} us), "
          f"resets/step/env {(c1-c0)/23/n:.5f}", flush=True)
    env.close()
