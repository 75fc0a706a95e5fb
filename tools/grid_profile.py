"""A few fused-rollout launches (for ncu): python tools/grid_profile.py [kind=1] [n_envs=262144] [K=32]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import torch
import neorl_industrial as ni
from neorl_industrial import _native as N
kind = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
K = int(sys.argv[3]) if len(sys.argv) > 3 else 32
env = ni.NativeEnv(kind, n, device=0, seed=0)
env.reset_device()
for _ in range(3):
    env.rollout_device(K, N.POLICY_UNIFORM)
torch.cuda.synchronize()
print("ok")
