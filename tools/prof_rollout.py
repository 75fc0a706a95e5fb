"""A few fused-rollout launches of ChemicalReactor-v0 (uniform policy, K = 64) in steady state, for ncu:
   ncu --set full --import-source on --clock-control none -k regex:rollout_kernel --launch-skip 12 --launch-count 1 \
       -o gpurun_out/prof python tools/prof_rollout.py [n_envs] [one_warp_ctas]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import torch
import neorl_industrial as ni
from neorl_industrial import _native as N
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
if len(sys.argv) > 2:
    os.environ["NIG_ROLLOUT_BLOCK"] = sys.argv[2]
env = ni.NativeEnv(N.ENV_CHEMICAL_REACTOR, n, device=0, seed=0)
env.reset_device()
for _ in range(14):
    env.rollout_device(64, N.POLICY_UNIFORM)
torch.cuda.synchronize()
print(env.stats_dict())
