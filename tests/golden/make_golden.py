"""Generate the golden fixtures by running the UNMODIFIED reference env code (teacher-forced).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
Reference version: _version.py -> 0.1.dev68+g9b77c89e2.d20250827; numpy 2.3.5 (NEP 50), fp32 actions,
fp32-representable noise (SURVEY.md section 8c "canonical for parity").

Fixtures (all small, committed):
  <env>_forced.npz : M independent (state, ep_step, action, noise) -> env.step() tuples, sampled to hit every
                     branch of the dynamics / reward / constraint / termination logic.
  <env>_trace.npz  : BASELINE config #1 -- one env, 1000 steps, random fp32 actions, reset on done; everything
                     needed to replay it teacher-forced (states, actions, noise, reset states, outputs).
  reactor_freerun.npz : 16 free-running 500-step episodes (actions + noise recorded) for the drift test.
  reactor_dataset_stats.json : summary statistics of the reference's get_dataset('mixed'/'expert').
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

ENV_CLASS = {"reactor": "ChemicalReactorEnv", "grid": "PowerGridEnv", "robot": "RobotAssemblyEnv"}
NOISE_SHAPES = {"reactor": [None, None], "grid": [8, 8, 7], "robot": []}
NOISE_SIGMA = {"reactor": [0.1, 500.0], "grid": [0.005, 1.0, 2.0], "robot": []}
NZ = {"reactor": 2, "grid": 23, "robot": 0}


def draw_noise(rng, name):
    """fp32-representable noise, flat [NZ], plus the list to push into the NoiseQueue."""
    parts, flat = [], []
    for shape, sig in zip(NOISE_SHAPES[name], NOISE_SIGMA[name]):
        if shape is None:
            v = np.float32(rng.normal(0, sig))
            parts.append(float(v)); flat.append(np.array([v], np.float32))
        else:
            v = rng.normal(0, sig, shape).astype(np.float32)
            parts.append(v.astype(np.float64)); flat.append(v)
    return parts, (np.concatenate(flat) if flat else np.zeros(0, np.float32))


def forced_step(env, q, state, ep_step, action, noise_parts):
    env.state = state.copy()
    env.current_step = int(ep_step)
    env.done = False
    env.violation_count = 0
    env.total_violations = 0
    q.queue.clear()
    q.push(*noise_parts)
    obs, reward, terminated, truncated, info = env.step(action.copy())
    sm = info["safety_metrics"]
    assert not q.queue
    return (obs.astype(np.float32), np.float64(reward), bool(terminated), bool(truncated),
            int(sm.violation_count), int(sm.critical_violations), bool(info["critical_shutdown"]),
            int(info["violations"]))


def viol_mask_of(env, state, action):
    a = np.clip(action, env.action_space.low, env.action_space.high)
    m = 0
    for k, c in enumerate(env.safety_constraints):
        if not c.check_fn(state, a):
            m |= 1 << k
    return m


# ------------------------------------------------------------------------------------------------
def sample_reactor_states(rng, env, m):
    S = np.zeros((m, 12), np.float32)
    steps = np.zeros(m, np.int32)
    for i in range(m):
        mode = i % 4
        if mode == 0:      # nominal: reset distribution
            S[i] = env._get_initial_state()
            steps[i] = rng.integers(0, 499)
        else:              # wide: every branch reachable
            S[i] = [rng.uniform(300, 360), rng.uniform(2.0e5, 5.3e5), rng.uniform(8, 102), rng.uniform(4, 52),
                    rng.uniform(0, 2) if rng.random() < 0.9 else 0.0, rng.uniform(49.9, 100), rng.uniform(285, 300),
                    0.0 if rng.random() < 0.5 else rng.uniform(0, 100), float(rng.random() < 0.3),
                    float(rng.random() < 0.4), rng.uniform(0, 100), rng.uniform(0, 51)]
            steps[i] = rng.choice([rng.integers(0, 497), 498, 499]) if mode == 3 else rng.integers(0, 499)
        if mode == 2:      # clamp edges hit exactly
            j = i % 12
            if j == 2: S[i, 3] = 5.0
            if j == 6: S[i, 3] = 50.0
            if j == 10: S[i, 2] = 10.0 if rng.random() < 0.5 else 100.0
            if i % 5 == 0: S[i, 10] = rng.choice([0.0, 100.0, 4.99, 95.01, 20.0, 90.0, 30.0, 80.0])
            if i % 7 == 0: S[i, 0] = rng.choice([350.0, 345.0, 340.0, 349.99, 350.01])
            if i % 11 == 0: S[i, 1] = rng.choice([506625.0, 480000.0, 506625.03, 101325.0])
    return S, steps


def sample_grid_states(rng, env, m):
    S = np.zeros((m, 32), np.float32)
    steps = np.zeros(m, np.int32)
    for i in range(m):
        S[i] = env._get_initial_state()
        mode = i % 4
        steps[i] = rng.integers(0, 999)
        if mode >= 1:
            S[i, 0] = rng.uniform(-1.2, 1.2) if mode == 1 else rng.normal(0, 0.3)
        if mode == 2:
            S[i, 1:9] = rng.uniform(0.88, 1.12, 8)
            S[i, 9:17] = rng.uniform(-1, 101, 8)
        if mode == 3:
            steps[i] = rng.choice([998, 999, rng.integers(0, 998)])
            k = rng.integers(0, 8)
            S[i, 1 + k] = rng.choice([0.95, 1.05, 0.9, 1.1, 0.94999, 1.05001])
            S[i, 9 + k] = rng.choice([0.0, 100.0, 99.5, 0.5])
            S[i, 17:25] = np.maximum(S[i, 17:25] + rng.normal(0, 30, 8), 0)
    return S, steps


def sample_robot_states(rng, env, m):
    S = np.zeros((m, 24), np.float32)
    steps = np.zeros(m, np.int32)
    for i in range(m):
        mode = i % 4
        steps[i] = rng.integers(0, 999) if mode != 3 else rng.choice([998, 999, 5])
        if mode == 0:
            S[i] = env._get_initial_state()
        else:
            q = rng.uniform(-np.pi, np.pi, 7) if mode == 1 else rng.uniform(-1.0, 1.0, 7)
            if mode == 2 and i % 8 == 2:
                # near the assembly target: search a configuration close to [0.3, 0, 0.4]
                best = None
                for _ in range(4000):
                    qq = rng.uniform(-np.pi, np.pi, 7)
                    p, _ = env._forward_kinematics(qq)
                    d = np.linalg.norm(p - env.target_position)
                    if best is None or d < best[0]:
                        best = (d, qq)
                q = best[1]
            if i % 9 == 0:
                q[rng.integers(0, 7)] = rng.choice([np.pi, -np.pi, 2.0, -2.0, 1.99, 3.1])
            p, quat = env._forward_kinematics(q)
            S[i, 0:3] = p + (rng.normal(0, 0.02, 3) if mode == 1 else 0)
            S[i, 3:7] = quat
            S[i, 7:14] = q
            S[i, 14:17] = rng.normal(0, 0.5, 3)
            if i % 6 == 0:
                S[i, 18:21] = rng.choice([0.0, 10.0, 49.9, 50.0, 79.0, 81.0, -60.0], 3)
            S[i, 21:24] = rng.uniform(0, 1, 3)
    return S, steps


SAMPLERS = {"reactor": sample_reactor_states, "grid": sample_grid_states, "robot": sample_robot_states}


def make_forced(envs, name, m, seed):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    env = getattr(envs, ENV_CLASS[name])()
    S, steps = SAMPLERS[name](rng, env, m)
    A = rng.uniform(-1.5, 1.5, (m, env.action_dim)).astype(np.float32)
    A[::13] = np.clip(A[::13], -1, 1)
    A[5::17] = 0.0
    out = {k: [] for k in ("next_state", "reward", "terminated", "truncated", "n_viol", "n_crit", "crit", "viol_mask", "noise")}
    with ref_loader.NoiseQueue() as q:
        for i in range(m):
            parts, flat = draw_noise(rng, name)
            vm = viol_mask_of(env, S[i], A[i])
            ns, r, te, tr, nv, nc, cs, ev = forced_step(env, q, S[i], steps[i], A[i], parts)
            assert ev == nv and bin(vm).count("1") == nv
            for k, v in zip(("next_state", "reward", "terminated", "truncated", "n_viol", "n_crit", "crit", "viol_mask", "noise"),
                            (ns, r, te, tr, nv, nc, cs, vm, flat)):
                out[k].append(v)
    np.savez_compressed(os.path.join(HERE, f"{name}_forced.npz"),
                        state=S, ep_step=steps, action=A,
                        noise=np.array(out["noise"], np.float32).reshape(m, NZ[name]),
                        next_state=np.array(out["next_state"], np.float32),
                        reward=np.array(out["reward"], np.float64),
                        terminated=np.array(out["terminated"], bool), truncated=np.array(out["truncated"], bool),
                        n_viol=np.array(out["n_viol"], np.int32), n_crit=np.array(out["n_crit"], np.int32),
                        crit=np.array(out["crit"], bool), viol_mask=np.array(out["viol_mask"], np.uint8))
    te, tr, cr = np.array(out["terminated"]), np.array(out["truncated"]), np.array(out["crit"])
    print(f"{name}_forced: {m} tuples, terminated {te.sum()}, truncated {tr.sum()}, critical {cr.sum()}, "
          f"violations {np.sum(out['n_viol'])}")


def make_trace(envs, name, n_steps, seed):
    """BASELINE config #1: performance_benchmark.py:106-133 loop, recorded for teacher-forced replay."""
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    env = getattr(envs, ENV_CLASS[name])()
    env.action_space.seed(seed)
    obs, _ = env.reset()
    rec = {k: [] for k in ("state", "action", "noise", "next_state", "reward", "terminated", "truncated", "n_viol",
                           "crit", "viol_mask", "ep_step", "reset_state", "ep_violations", "total_violations")}
    with ref_loader.NoiseQueue() as q:
        for t in range(n_steps):
            a = env.action_space.sample()
            parts, flat = draw_noise(rng, name)
            q.queue.clear(); q.push(*parts)
            s0 = env.state.copy(); st0 = env.current_step
            vm = viol_mask_of(env, s0, a)
            obs, r, te, tr, info = env.step(a)
            rec["state"].append(s0); rec["action"].append(a); rec["noise"].append(flat)
            rec["next_state"].append(obs.copy()); rec["reward"].append(np.float64(r))
            rec["terminated"].append(te); rec["truncated"].append(tr)
            rec["n_viol"].append(info["safety_metrics"].violation_count); rec["crit"].append(info["critical_shutdown"])
            rec["viol_mask"].append(vm); rec["ep_step"].append(st0)
            rec["ep_violations"].append(info["violations"]); rec["total_violations"].append(info["total_violations"])
            if te or tr:
                np.random.normal = q._orig          # reset draws come from the real global RNG
                obs, _ = env.reset()
                np.random.normal = q._normal
                rec["reset_state"].append(env.state.copy())
            else:
                rec["reset_state"].append(np.zeros(env.state_dim, np.float32))
    np.savez_compressed(os.path.join(HERE, f"{name}_trace.npz"),
                        state=np.array(rec["state"], np.float32), action=np.array(rec["action"], np.float32),
                        noise=np.array(rec["noise"], np.float32).reshape(n_steps, NZ[name]),
                        next_state=np.array(rec["next_state"], np.float32), reward=np.array(rec["reward"], np.float64),
                        terminated=np.array(rec["terminated"], bool), truncated=np.array(rec["truncated"], bool),
                        n_viol=np.array(rec["n_viol"], np.int32), crit=np.array(rec["crit"], bool),
                        viol_mask=np.array(rec["viol_mask"], np.uint8), ep_step=np.array(rec["ep_step"], np.int32),
                        reset_state=np.array(rec["reset_state"], np.float32),
                        ep_violations=np.array(rec["ep_violations"], np.int32),
                        total_violations=np.array(rec["total_violations"], np.int32))
    d = np.array(rec["terminated"]) | np.array(rec["truncated"])
    print(f"{name}_trace: {n_steps} steps, {d.sum()} episodes ended, {np.sum(rec['n_viol'])} violations")


def make_freerun(envs, n_ep, seed):
    rng = np.random.default_rng(seed)
    np.random.seed(seed)
    env = envs.ChemicalReactorEnv()
    T = 500
    init = np.zeros((n_ep, 12), np.float32)
    A = rng.uniform(-1, 1, (n_ep, T, 3)).astype(np.float32)
    A[n_ep // 2:] *= 0.3      # gentler actions -> longer episodes
    NZv = np.zeros((n_ep, T, 2), np.float32)
    states = np.zeros((n_ep, T, 12), np.float32)
    rewards = np.zeros((n_ep, T), np.float64)
    flags = np.zeros((n_ep, T), np.uint8)
    length = np.zeros(n_ep, np.int32)
    with ref_loader.NoiseQueue() as q:
        for e in range(n_ep):
            np.random.normal = q._orig
            env.reset()
            np.random.normal = q._normal
            init[e] = env.state
            for t in range(T):
                parts, flat = draw_noise(rng, "reactor")
                NZv[e, t] = flat
                q.queue.clear(); q.push(*parts)
                obs, r, te, tr, info = env.step(A[e, t])
                states[e, t] = obs; rewards[e, t] = r
                flags[e, t] = (1 if te else 0) | (2 if tr else 0) | (4 if info["critical_shutdown"] else 0)
                length[e] = t + 1
                if te or tr:
                    break
    np.savez_compressed(os.path.join(HERE, "reactor_freerun.npz"), init=init, action=A, noise=NZv,
                        states=states, rewards=rewards, flags=flags, length=length)
    print("reactor_freerun: lengths", length.tolist())


def make_dataset_stats(envs, seed):
    out = {}
    for quality in ("mixed", "expert", "medium", "random"):
        np.random.seed(seed)
        env = envs.ChemicalReactorEnv()
        d = env.get_dataset(quality)
        term = d["terminals"]
        ends = np.flatnonzero(term)
        out[quality] = {
            "n": int(len(d["rewards"])), "n_terminals": int(term.sum()),
            "obs_dtype": str(d["observations"].dtype), "act_dtype": str(d["actions"].dtype),
            "rew_dtype": str(d["rewards"].dtype), "term_dtype": str(d["terminals"].dtype),
            "keys": sorted(d.keys()),
            "reward_mean": float(d["rewards"].mean()), "reward_std": float(d["rewards"].std()),
            "reward_median": float(np.median(d["rewards"])),
            "action_mean": d["actions"].mean(0).tolist(), "action_std": d["actions"].std(0).tolist(),
            "obs_mean": d["observations"].mean(0).tolist(), "obs_std": d["observations"].std(0).tolist(),
            "frac_action_saturated": float(np.mean(np.abs(d["actions"]) >= 1.0)),
            "timeouts_any": bool(d["timeouts"].any()),
        }
        print(quality, {k: out[quality][k] for k in ("n", "n_terminals", "reward_mean", "reward_std")})
    with open(os.path.join(HERE, "reactor_dataset_stats.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    envs = ref_loader.load_reference_envs()
    which = sys.argv[1:] or ["forced", "trace", "freerun", "dataset"]
    if "forced" in which:
        make_forced(envs, "reactor", 6000, 1)
        make_forced(envs, "grid", 3000, 2)
        make_forced(envs, "robot", 2000, 3)
    if "trace" in which:
        make_trace(envs, "reactor", 1000, 0)
        make_trace(envs, "grid", 1000, 0)
        make_trace(envs, "robot", 1000, 0)
    if "freerun" in which:
        make_freerun(envs, 16, 5)
    if "dataset" in which:
        make_dataset_stats(envs, 0)
