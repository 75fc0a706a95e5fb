import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
from oracle import oracle as O
n = 256
b = ni.NativeEnv(0, n, device=0, seed=33, max_episode_steps=10)
orc = O.OracleEnv(0, n, seed=33, exp_mode=1, max_episode_steps=10)
b.reset_host(); orc.reset()
run = np.zeros(n, np.float32); gt = 0.0
def ostep(act):
    global run, gt
    _, r, fl, _ = orc.step(act, want_next_obs=False)
    run = (run + r).astype(np.float32)
    done = (fl & 3) > 0
    gt += float(run[done].astype(np.float64).sum()); run[done] = 0
for how, k in [("rollout", 3), ("step", 2), ("rollout", 4), ("step", 3), ("rollout", 2), ("step", 1), ("step", 1), ("rollout", 9)]:
    for _ in range(k):
        act = O.policy_actions(orc, O.POLICY_UNIFORM)
        if how == "step": b.step_host(act, want_obs=False)
        ostep(act)
    if how == "rollout": b.rollout_device(k, N.POLICY_UNIFORM)
    torch.cuda.synchronize()
    c, f = b.read_stats()
    print(how, k, "episodes", c[1], "gpu return_sum", f[0], "truth", gt, "diff", f[0] - gt, flush=True)
