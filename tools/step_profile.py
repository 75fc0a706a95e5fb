"""Launch the single-step kernel a few times per env kind (for ncu): python tools/step_profile.py [launches]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import torch
import neorl_industrial as ni
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda", 0)
for kind, n in ((0, 1 << 24), (1, 1 << 20), (2, 1 << 20)):
    env = ni.NativeEnv(kind, n, device=0, seed=0)
    env.reset_device()
    acts = torch.rand((env.A, env.pitch), device=dev) * 2 - 1
    rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
    for _ in range(reps):
        env.step_device(acts, reward=rew, flags=fl, viol_mask=vm)
    torch.cuda.synchronize()
    env.close()
print("ok")
