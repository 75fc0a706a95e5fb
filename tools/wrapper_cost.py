"""Cost of SafetyWrapper bound constraints in the fused reactor rollout: plain env vs +1 / +2 / +4 declarative bounds
(65,536 envs, 1,024 steps, K = 64, env slices on streams) -> env-steps/s."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
from neorl_industrial.safety import BoundConstraint, SafetyWrapper

BANDS = [BoundConstraint("temperature_band", 0, 280.0, 330.0, penalty=-100.0),
         BoundConstraint("pressure_band", 1, 101325.0, 400000.0, penalty=-100.0),
         BoundConstraint("coolant_band", 2, 0.0, 100.0, penalty=-10.0),
         BoundConstraint("feed_band", 3, 0.0, 100.0, penalty=-10.0)]
n, T = 65536, 1024
for extra in (0, 1, 2, 4):
    env = ni.make("ChemicalReactor-v0", num_envs=n, device="cuda:0", seed=0)
    if extra:
        env = SafetyWrapper(env, constraints=BANDS[:extra])
    nat = env.native
    nat.reset_device()
    for _ in range(2): nat.rollout_steps_device(T, 64, N.POLICY_UNIFORM)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); nat.rollout_steps_device(T, 64, N.POLICY_UNIFORM); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    print(f"extra bound constraints = {extra}: {n * T / ms / 1e-3:.4g} env-steps/s ({ms:.3f} ms / {T} steps)", flush=True)
    nat.close()
