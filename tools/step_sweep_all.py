"""Single-step kernel: persistent TMA-pipelined vs one-tile-per-CTA, all envs -> algorithmic GB/s and fraction of the measured HBM peak."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
dev = torch.device("cuda", 0)
BYTES = {0: 122, 1: 302, 2: 234}
def run(kind, n, pipe, reps=20):
    os.environ["NIG_STEP_PIPE"] = str(pipe)
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
which = sys.argv[1:] or ["reactor", "grid", "robot"]
for kind, name in ((0, "reactor"), (1, "grid"), (2, "robot")):
    if name not in which:
        continue
    for n in ((1 << 20, 1 << 22, 1 << 24) if kind == 0 else (1 << 20, 1 << 22)):
        for pipe, pvec in ((0, 0), (1, 2)) if kind == 0 else ((0, 0),):
            os.environ["NIG_STEP_PIPE_VEC"] = str(pvec)
            gbs, ms = run(kind, n, pipe)
            print(f"{name:8s} n={n:9d} pipe={pipe} pipe_vec={pvec}: {gbs:7.0f} GB/s  ({ms*1e3:8.1f} us)  frac {gbs/6450.9:.3f}", flush=True)
