"""PowerGrid / RobotAssembly single step: one-tile-per-CTA step_kernel vs the persistent TMA-pipelined step_pipe_kernel
(how the routing in csrc/nig_step.cu was decided; needs the experiment switch NIG_STEP_PIPE_ALL compiled back in), steady-state episode mix, 1M and 4M envs -> us per launch, fraction of the measured HBM peak."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
dev = torch.device("cuda", 0)
BYTES = {1: 302, 2: 234}
for kind, name in ((1, "grid"), (2, "robot")):
    for n in (1 << 20, 1 << 22):
        res = {}
        for pipe in (0, 1, 2, 0, 1, 2):
            if pipe: os.environ["NIG_STEP_PIPE_ALL"] = str(pipe)
            else: os.environ.pop("NIG_STEP_PIPE_ALL", None)
            env = ni.NativeEnv(kind, n, device=0, seed=0); env.reset_device()
            env.rollout_device(64, N.POLICY_UNIFORM)
            acts = torch.rand((env.A, env.pitch), device=dev) * 2 - 1
            rew, fl, vm = env.empty(), env.empty(dtype=torch.uint8), env.empty(dtype=torch.uint8)
            for _ in range(3): env.step_device(acts, reward=rew, flags=fl, viol_mask=vm)
            torch.cuda.synchronize()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
            for a, b in evs:
                a.record(); env.step_device(acts, reward=rew, flags=fl, viol_mask=vm); b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
            st = env.get_state_host()[0]
            res.setdefault(pipe, []).append((ms, int(st.view(np.uint32).astype(np.uint64).sum())))
            env.close()
        for pipe in (0, 1, 2):
            ms = min(x[0] for x in res[pipe])
            print(f"{name} n={n} pipe={pipe}: {ms * 1e3:.1f} us, {BYTES[kind] * n / ms / 1e6:.0f} GB/s ({BYTES[kind] * n / ms / 1e6 / 6450.9:.3f} of HBM peak) checksum {res[pipe][0][1]}", flush=True)
