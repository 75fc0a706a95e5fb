"""Single-step kernel sweep: vector width x env count -> achieved algorithmic GB/s (122 B / env-step)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
dev = torch.device("cuda", 0)
BYTES = {0: 122, 1: 302, 2: 234}
def run(kind, n, vec, reps=20, lean=True):
    os.environ["NIG_STEP_VEC"] = str(vec)
    env = ni.NativeEnv(kind, n, device=0, seed=0)
    env.reset_device()
    acts = torch.rand((env.A, env.pitch), device=dev) * 2 - 1
    rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
    for _ in range(3): env.step_device(acts, reward=rew, flags=fl, viol_mask=vm)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in evs:
        a.record(); env.step_device(acts, reward=rew, flags=fl, viol_mask=vm); b.record()
    torch.cuda.synchronize()
    ms = np.median([a.elapsed_time(b) for a, b in evs])
    env.close()
    return BYTES[kind] * n / (ms * 1e-3) / 1e9, ms
for kind, name in ((0, "reactor"), (1, "grid"), (2, "robot")):
    for n in ((1 << 20, 1 << 22, 1 << 24) if kind == 0 else (1 << 20, 1 << 22)):
        for vec in ((1, 2, 4) if kind == 0 else (1, 2)):
            gbs, ms = run(kind, n, vec)
            print(f"{name:8s} n={n:9d} vec={vec}: {gbs:7.0f} GB/s  ({ms*1e3:8.1f} us)  frac {gbs/6450.9:.3f}", flush=True)
