"""PowerGrid-v0 single step at 1M envs, steady-state episode mix: time per launch, fraction of the HBM peak (302 B/env-step).
   NIG_GRID_STEP=0 -> the generic one-tile kernel; argv: [n_envs] [track_returns 0/1]"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
track = int(sys.argv[2]) if len(sys.argv) > 2 else 1
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
dev = torch.device("cuda", 0)
env = ni.NativeEnv(N.ENV_POWER_GRID, n, device=0, seed=0)
env.track_returns(bool(track))
env.reset_device()
env.rollout_device(64, N.POLICY_UNIFORM)
acts = torch.rand((env.A, env.pitch), device=dev) * 2 - 1
rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5): env.step_device(acts, reward=rew, flags=fl, viol_mask=vm)
torch.cuda.synchronize()
ts = []
for _ in range(30):
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); env.step_device(acts, reward=rew, flags=fl, viol_mask=vm); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
us = float(np.median(ts)) * 1e3
byt = (302 + (16 if track else 0)) * n
print(f"NIG_GRID_STEP={os.environ.get('NIG_GRID_STEP', '1')} n={n} track_returns={track}: {us:.1f} us/launch, {n / us * 1e6:.4g} env-steps/s, "
      f"{byt / us * 1e-3:.0f} GB/s algorithmic ({302 + (16 if track else 0)} B/env-step)", flush=True)
