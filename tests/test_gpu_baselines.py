"""In-kernel baseline controllers (benchmarks/baseline_agents.py) vs the host agents (the reference's numpy fp64
arithmetic) driving the CPU oracle env: same trajectories bit for bit."""
import numpy as np
import pytest

from util import KINDS, assert_bits_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import neorl_industrial as ni
    from neorl_industrial import _native as N
    from neorl_industrial.benchmarks import BaselineAgentFactory
    from oracle import oracle as O
    return ni, N, O, BaselineAgentFactory


def _host_actions(agents, obs):
    return np.stack([ag.act(o) for ag, o in zip(agents, obs)]).astype(np.float32)


@pytest.mark.parametrize("name", ["reactor", "grid", "robot"])
@pytest.mark.parametrize("kind", ["pid", "mpc", "constant"])
def test_baseline_controller_rollout_vs_host_agent(mods, name, kind):
    ni, N, O, Factory = mods
    ek, n, T, K = KINDS[name], 96, 90, 32
    cls = {"reactor": ni.ChemicalReactorEnv, "grid": ni.PowerGridEnv, "robot": ni.RobotAssemblyEnv}[name]
    env = cls(num_envs=n, seed=5)
    S, A = env.state_dim, env.action_dim
    kw = {}
    if kind == "pid":       # a setpoint near the operating point so that the controller is not permanently saturated
        sp = {"reactor": [320.0, 253312.0, 50.0], "grid": [0.0] + [1.0] * 7, "robot": [0.3, 0.0, 0.4, 0, 0, 0, 1.0]}[name]
        kw = {"kp": 0.05, "ki": 0.001, "kd": 0.02, "setpoint": np.array(sp, np.float64)}
    elif kind == "constant":
        kw = {"constant_action": np.linspace(-0.4, 0.6, A)}
    dev_agent = Factory.create(kind, S, A, **kw)
    host_agents = [Factory.create(kind, S, A, **kw) for _ in range(n)]      # one agent object per env, like upstream
    orc = O.OracleEnv(ek, n, auto_reset=True, seed=5, exp_mode=1)
    orc.reset()
    res = env.rollout(T, dev_agent, steps_per_launch=K, reset=True)
    viol = np.zeros(n, np.int64)
    for t in range(T):
        a = _host_actions(host_agents, orc.state)
        _, r, fl, vm = orc.step(a, want_next_obs=False)
        viol += np.array([bin(int(v)).count("1") for v in vm])
    assert_bits_equal(res["obs"], orc.state, f"{kind} final state")
    assert np.array_equal(res["violations"], viol)
    # the device-resident PID state carries over to the next call; a reset of it equals a new agent
    if kind == "pid":
        res2 = env.rollout(7, dev_agent, steps_per_launch=K)
        for t in range(7):
            orc.step(_host_actions(host_agents, orc.state), want_next_obs=False)
        assert_bits_equal(res2["obs"], orc.state, "pid state carried across calls")
        env.native.reset_policy_state()
        host_agents = [Factory.create(kind, S, A, **kw) for _ in range(n)]
        res3 = env.rollout(5, dev_agent, steps_per_launch=K)
        for t in range(5):
            orc.step(_host_actions(host_agents, orc.state), want_next_obs=False)
        assert_bits_equal(res3["obs"], orc.state, "pid after reset_policy_state")
    env.close()


def test_random_baseline_is_in_range(mods):
    ni, N, O, Factory = mods
    env = ni.ChemicalReactorEnv(num_envs=512, seed=3)
    ag = Factory.create("random", 12, 3, action_low=-0.25, action_high=0.5)
    res = env.rollout(64, ag, reset=True)
    assert res["stats"]["steps"] == 512 * 64 and np.isfinite(res["reward_sum"]).all()
    env.close()
