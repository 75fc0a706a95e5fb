"""A few PowerGrid fused-rollout and single-step launches (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import torch
import neorl_industrial as ni
from neorl_industrial import _native as N
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 1
env = ni.NativeEnv(kind, 1 << 18, device=0, seed=0)
env.reset_device()
for _ in range(3):
    env.rollout_device(32, N.POLICY_UNIFORM)
torch.cuda.synchronize()
print("ok")
