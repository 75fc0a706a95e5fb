"""Headline workload at different K (fused steps per launch): what the launch boundaries of K = 64 cost.
   python tools/k_sweep.py [n_envs]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tools")]
from ab_rollout import run
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
for K in (32, 64, 128, 250, 500, 1000):
    med, best, ret = run(n, K=K)
    print(f"n={n} K={K:5d} slices: median {med:.4g}  best {best:.4g} env-steps/s", flush=True)
