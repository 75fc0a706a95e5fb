import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import neorl_industrial as ni
from neorl_industrial import _native as N
from neorl_industrial.vector import make_constraint
from oracle import oracle as O
g = np.load(os.path.join(ROOT, "tests/golden/robot_forced.npz"))
m = len(g["reward"])
def run(cons=None, n=m):
    env = ni.NativeEnv(2, n, device=0, auto_reset=False, constraints=cons)
    env.set_state_host(g["state"][:n], g["ep_step"][:n], np.zeros(n, np.int32), np.zeros(n, np.uint8))
    obs, nxt, r, fl, vm = env.step_host(g["action"][:n], want_next_obs=True)
    return nxt.copy(), r.copy(), fl.copy(), vm.copy()
nxt, r, fl, vm = run()
gt = g["terminated"]
t = (fl & 1) > 0
print("gpu T gold F:", (t & ~gt).sum(), " gpu F gold T:", (~t & gt).sum(), " trunc mismatch", (((fl & 2) > 0) != g["truncated"]).sum())
orc_done = O.is_done(2, nxt)
print("oracle is_done on GPU next_obs vs gpu term (non-crit rows):", ((orc_done != t) & ~g["crit"]).sum())
print("oracle is_done vs gold term (non-crit):", ((orc_done != gt) & ~g["crit"]).sum())
bad = np.flatnonzero(t != gt)
print("bad rows mod 32:", np.bincount(bad % 32, minlength=32))
print("bad rows first 40:", bad[:40])
print("flags histogram", np.bincount(fl, minlength=16)[:16])
cons = [make_constraint(N.CON_BUILTIN, cid=k, penalty=p, critical=c) for k, (p, c) in enumerate([(-100.5, 1), (-200, 1), (-50, 0)])]
nxt2, r2, fl2, vm2 = run(cons)
print("non-default-cons path: term mismatch", (((fl2 & 1) > 0) != gt).sum())
nxt3, r3, fl3, vm3 = run(None, 1)
print("single env row0:", fl3, g["terminated"][0])
