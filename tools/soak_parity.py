"""Soak parity run (not part of the test suite): large random populations, many steps, every env and in-kernel policy,
fused rollout + single steps + explicit resets interleaved, CUDA path vs the CPU oracle bit for bit.
    python tools/soak_parity.py [n_envs] [rounds]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200"), os.path.join(ROOT, "tests")]
import numpy as np, torch
import neorl_industrial as ni
from neorl_industrial import _native as N
from oracle import oracle as O
from util import assert_bits_equal

n = int(sys.argv[1]) if len(sys.argv) > 1 else 150_001
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
threads = max(1, len(os.sched_getaffinity(0)))
total = 0
t0 = time.time()
for kind, name, cls in ((0, "reactor", ni.ChemicalReactorEnv), (1, "grid", ni.PowerGridEnv), (2, "robot", ni.RobotAssemblyEnv)):
    for seed in (1, 2):
        env = ni.NativeEnv(kind, n, device=0, seed=seed, env_id_offset=seed * 1000)
        orc = O.OracleEnv(kind, n, auto_reset=True, seed=seed, env_id0=seed * 1000, exp_mode=1, threads=threads)
        assert_bits_equal(env.reset_host(), orc.reset(), "reset")
        _, _, _, pp = cls.dataset_policy("mixed")
        opp = O.copy_policy_params(pp)
        rng = np.random.default_rng(seed)
        dev = env.torch_device()
        for r in range(rounds):
            K = int(rng.integers(20, 90))
            pol = (N.POLICY_UNIFORM, N.POLICY_PCTRL, N.POLICY_ZERO)[r % 3]
            env.rollout_device(K, pol, params=pp if pol == N.POLICY_PCTRL else None)
            O.rollout(orc, K, {N.POLICY_UNIFORM: O.POLICY_UNIFORM, N.POLICY_PCTRL: O.POLICY_PCTRL, N.POLICY_ZERO: O.POLICY_ZERO}[pol],
                      pp=opp if pol == N.POLICY_PCTRL else None)
            for _ in range(3):                                   # a few single steps with host actions
                a = rng.uniform(-1.5, 1.5, (n, env.A)).astype(np.float32)
                obs, _, rew, fl, vm = env.step_host(a)
                _, o_r, o_fl, o_vm = orc.step(a, want_next_obs=False)
                assert_bits_equal(fl, o_fl, f"{name} flags"); assert_bits_equal(rew, o_r, f"{name} reward")
                assert_bits_equal(vm, o_vm, f"{name} viol"); assert_bits_equal(obs, orc.state, f"{name} state")
            if r % 2 == 1:
                mask = (rng.random(n) < 0.3).astype(np.uint8)
                assert_bits_equal(env.reset_host(mask=mask), orc.reset(mask=mask), "masked reset")
            st, es, ev, dn = env.get_state_host()
            assert_bits_equal(st, orc.state, f"{name} seed {seed} round {r} state")
            assert np.array_equal(es, orc.ep_step) and np.array_equal(ev, orc.ep_viol)
            total += n * (K + 3)
        c, _ = env.read_stats()
        assert c[:6].tolist() == orc.stats[:6].tolist(), (c[:6], orc.stats[:6])
        print(f"{name} seed {seed}: ok, episodes {int(c[1])}, violations {int(c[5])}, critical {int(c[4])}", flush=True)
        env.close()
# ---- constraint descriptors: default prefix + extra bounds (CONS_PREFIX) and a re-ordered / partial set (CONS_GENERIC)
from neorl_industrial.vector import make_constraint
f32 = lambda x: float(np.float32(x))
for kind, name in ((0, "reactor"), (1, "grid"), (2, "robot")):
    spec = N.env_spec(kind)
    builtins = [make_constraint(N.CON_BUILTIN, cid=k, penalty=spec.constraints[k].penalty, critical=bool(spec.constraints[k].critical)) for k in range(3)]
    extras = [make_constraint(N.CON_BOUND, si=0, ai=0, coef=f32(0.1), lo=f32(-0.4 if kind else 300.0), hi=f32(0.4 if kind else 325.0), penalty=-40.0),
              make_constraint(N.CON_BOUND, si=1, lo=f32(0.97 if kind == 1 else (-0.3 if kind == 2 else 1.2e5)), hi=f32(1.03 if kind == 1 else (0.3 if kind == 2 else 3.5e5)), penalty=-15.0, critical=True)]
    for tag, cons in (("prefix", builtins + extras), ("generic", [extras[1], builtins[2], extras[0], builtins[0]])):
        env = ni.NativeEnv(kind, n, device=0, seed=7, constraints=cons)
        ocons = [O.Con(c.kind, c.id, c.si, c.ai, c.coef, c.lo, c.hi, c.penalty, c.critical) for c in cons]
        orc = O.OracleEnv(kind, n, auto_reset=True, seed=7, exp_mode=1, threads=threads, builtin=False, extra_cons=ocons)
        assert_bits_equal(env.reset_host(), orc.reset(), "reset")
        rng = np.random.default_rng(5)
        for r in range(3):
            K = int(rng.integers(20, 70))
            env.rollout_device(K, N.POLICY_UNIFORM); O.rollout(orc, K, O.POLICY_UNIFORM)
            a = rng.uniform(-1.5, 1.5, (n, env.A)).astype(np.float32)
            obs, _, rew, fl, vm = env.step_host(a)
            _, o_r, o_fl, o_vm = orc.step(a, want_next_obs=False)
            assert_bits_equal(fl, o_fl, f"{name} {tag} flags"); assert_bits_equal(rew, o_r, f"{name} {tag} reward")
            assert_bits_equal(vm, o_vm, f"{name} {tag} viol"); assert_bits_equal(obs, orc.state, f"{name} {tag} state")
            total += n * (K + 1)
        c, _ = env.read_stats()
        assert c[:6].tolist() == orc.stats[:6].tolist() and c[8:8 + len(cons)].tolist() == orc.stats[8:8 + len(cons)].tolist()
        print(f"{name} constraints {tag}: ok, per-constraint violations {c[8:8 + len(cons)].tolist()}", flush=True)
        env.close()
print(f"soak parity OK: {total:.3e} env-steps compared bit for bit in {time.time() - t0:.0f} s ({n} envs, {rounds} rounds, {threads} oracle threads)")
