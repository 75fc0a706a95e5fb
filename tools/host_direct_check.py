"""nig_rollout_host: direct mode (the slices' kernels read / write the pinned host arrays: NIG_HOST_DIRECT bit 0 = ingest, bit 1 =
export) vs staged copies over populations: bit-identical outputs + wall clock per call.
   python tools/host_direct_check.py [n_envs ...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N

ns = [int(x) for x in sys.argv[1:]] or [65536, 131072, 262144, 524288, (1 << 20) + 333]
for n in ns:
    for kind, name in ((N.ENV_CHEMICAL_REACTOR, "reactor"), (N.ENV_POWER_GRID, "grid")):
        res = {}
        for direct in ("0", "1", "2", "3"):
            os.environ["NIG_HOST_DIRECT"] = direct
            os.environ["NIG_HOST_DIRECT_MAX_MB"] = "1e9,1e9"          # no size thresholds: the raw comparison
            env = ni.NativeEnv(kind, n, device=0, seed=5)
            init = env.pinned("init_states", (n, env.S), np.float32)
            init[:] = env.reset_host()
            ts = []
            for _ in range(8):
                t0 = time.perf_counter(); out = env.rollout_host(200, N.POLICY_UNIFORM, steps_per_launch=64, init_states=init); ts.append(time.perf_counter() - t0)
            res[direct] = {k: np.array(v, copy=True) for k, v in out.items()}
            res[direct]["ms"] = float(np.median(ts[2:])) * 1e3
            env.close()
        same = all(np.array_equal(res[d][k].view(np.uint8), res["0"][k].view(np.uint8)) for d in "123" for k in ("reward_sum", "violations", "episodes", "obs", "counters"))
        print(f"{name:8s} n={n:8d} 200 steps  staged {res['0']['ms']:.3f}  direct-in {res['1']['ms']:.3f}  direct-out {res['2']['ms']:.3f}  both {res['3']['ms']:.3f} ms  "
              f"bit-identical: {same}", flush=True)
        assert same
