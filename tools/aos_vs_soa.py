"""Single-step kernel: SoA device actions (plain path) vs the AoS layouts of the host / torch APIs, large population."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
dev = torch.device("cuda", 0)
def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3
for kind, A, S in ((0, 3, 12), (1, 8, 32)):
    n = 1 << 22
    env = ni.NativeEnv(kind, n, device=0, seed=0); env.reset_device()
    soa = torch.rand((A, env.pitch), device=dev) * 2 - 1
    aos = torch.rand((n, A), device=dev) * 2 - 1
    obs = torch.empty((n, S), device=dev); nobs = torch.empty((n, S), device=dev)
    rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
    a = t(lambda: env.step_device(soa, reward=rew, flags=fl, viol_mask=vm))
    b = t(lambda: env.step_device(aos, reward=rew, flags=fl, viol_mask=vm, action_layout=N.LAYOUT_AOS))
    c = t(lambda: env.step_device(aos, obs=obs, next_obs=nobs, reward=rew, flags=fl, viol_mask=vm, action_layout=N.LAYOUT_AOS, aux_layout=N.LAYOUT_AOS))
    print(f"kind {kind} n={n}: SoA plain {a:.1f} us | AoS actions {b:.1f} us | AoS actions + obs + next_obs out {c:.1f} us", flush=True)
    env.close()
