"""nig_rollout_host / IndustrialEnv.rollout: the fused rollout with HOST buffers (what bench.py's e2e times).
Checked against the CPU oracle (bit-exact) and, teacher-forced, against the reference's own free-running episodes."""
import os

import numpy as np
import pytest

from util import KINDS, assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import neorl_industrial as ni
    from neorl_industrial import _native as N
    from oracle import oracle as O
    return ni, N, O


def _chunked_sum(r_steps, K):
    """fp32 sum of per-step rewards the way the device does it: sequential inside a launch, launches added in order."""
    T, n = r_steps.shape
    total = np.zeros(n, np.float32)
    for c0 in range(0, T, K):
        part = np.zeros(n, np.float32)
        for t in range(c0, min(c0 + K, T)):
            part = (part + r_steps[t]).astype(np.float32)
        total = part if c0 == 0 else (total + part).astype(np.float32)
    return total


@pytest.mark.parametrize("name", ["reactor", "grid", "robot"])
def test_rollout_host_random_policy_vs_oracle(mods, name):
    ni, N, O = mods
    kind, n, T, K = KINDS[name], 1500, 150, 64
    cls = {"reactor": ni.ChemicalReactorEnv, "grid": ni.PowerGridEnv, "robot": ni.RobotAssemblyEnv}[name]
    env = cls(num_envs=n, seed=21, env_id_offset=64)
    orc = O.OracleEnv(kind, n, auto_reset=True, seed=21, env_id0=64, exp_mode=1)
    orc.reset()
    res = env.rollout(T, "random", steps_per_launch=K, reset=True)
    r_steps = np.zeros((T, n), np.float32); viol = np.zeros(n, np.int64); eps = np.zeros(n, np.int64)
    for t in range(T):
        a = O.policy_actions(orc, O.POLICY_UNIFORM)
        _, r, fl, vm = orc.step(a, want_next_obs=False)
        r_steps[t] = r
        viol += np.array([bin(int(v)).count("1") for v in vm])
        eps += (fl & (N.F_TERMINATED | N.F_TRUNCATED)) != 0
    assert_bits_equal(res["obs"], orc.state, "final state")
    assert np.array_equal(res["violations"], viol) and np.array_equal(res["episodes"], eps)
    assert_bits_equal(res["reward_sum"], _chunked_sum(r_steps, K), "reward_sum")
    st = res["stats"]
    assert st["steps"] == n * T and st["episodes"] == int(eps.sum()) and st["violations"] == int(viol.sum())
    # a second call continues where the first one stopped (no reset)
    res2 = env.rollout(10, "random", steps_per_launch=K)
    for t in range(10):
        orc.step(O.policy_actions(orc, O.POLICY_UNIFORM), want_next_obs=False)
    assert_bits_equal(res2["obs"], orc.state, "continued state")
    env.close()


def test_rollout_host_teacher_forced_vs_reference_episodes(mods, golden_dir):
    """16 free-running 500-step reference episodes (initial state, actions and noise from the unmodified reference)
    replayed through the fused kernel with teacher-forced actions AND noise, in 64-step launches."""
    ni, N, O = mods
    g = np.load(os.path.join(golden_dir, "reactor_freerun.npz"))
    n, T = g["action"].shape[:2]
    env = ni.ChemicalReactorEnv(num_envs=n, auto_reset=False, batched=True)
    orc = O.OracleEnv(O.REACTOR, n, auto_reset=False, exp_mode=1)
    orc.reset(init_states=g["init"])
    acts = np.ascontiguousarray(g["action"].transpose(1, 0, 2))      # [T, n, A]
    nz = np.ascontiguousarray(g["noise"].transpose(1, 0, 2))
    res = env.rollout(T, "actions", actions=acts, noise=nz, init_states=g["init"], steps_per_launch=64)
    r_steps = np.zeros((T, n), np.float32)
    for t in range(T):
        _, r, fl, vm = orc.step(acts[t], noise=nz[t], want_next_obs=False)
        r_steps[t] = r
    assert_bits_equal(res["obs"], orc.state, "final state vs oracle")
    assert_bits_equal(res["reward_sum"], _chunked_sum(r_steps, 64), "reward_sum vs oracle")
    # vs the reference itself: episode ends identical, final states within the stated drift tolerance (1e-3 relative)
    ended = g["length"] < T
    assert np.array_equal(res["episodes"].astype(bool), np.ones(n, bool))        # every episode ends within 500 steps
    last = g["states"][np.arange(n), g["length"] - 1]
    rel = np.abs(res["obs"] - last) / np.maximum(np.abs(last), 1e-3)
    assert rel.max() <= 1e-3, rel.max()
    assert ended.sum() >= 0
    env.close()


def test_rollout_host_argument_errors(mods):
    ni, N, O = mods
    env = ni.ChemicalReactorEnv(num_envs=8)
    with pytest.raises(ValueError):
        env.rollout(10, "nonsense")
    with pytest.raises(ValueError):
        env.rollout(0, "random")
    with pytest.raises(ValueError):
        env.rollout(4, "actions")
    env.close()
