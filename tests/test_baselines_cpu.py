"""Host side of the baseline controllers (no GPU): reference semantics of act() and the device descriptors."""
import numpy as np
import pytest

from neorl_industrial import _native as N
from neorl_industrial.benchmarks import BaselineAgentFactory, PIDControllerAgent


def test_factory_and_errors():
    for k in ("random", "pid", "mpc", "constant"):
        ag = BaselineAgentFactory.create(k, 12, 3)
        a = np.asarray(ag.act(np.linspace(-2, 2, 12).astype(np.float32)))
        assert a.shape == (3,) and np.all(np.abs(a) <= 1.0)
        pid, pp = ag.device_policy()
        assert pid == N.POLICY_BASELINE and pp.baseline.kind == {"random": 0, "pid": 1, "mpc": 2, "constant": 3}[k]
    with pytest.raises(ValueError, match="Unknown baseline agent type"):
        BaselineAgentFactory.create("dqn", 12, 3)


def test_pid_matches_reference_recurrence():
    """baseline_agents.py:62-81: e = sp - pv; I += e; a = clip(kp e + ki I + kd (e - e_prev))."""
    ag = PIDControllerAgent(4, 2, kp=0.5, ki=0.1, kd=0.05, setpoint=np.array([1.0, -1.0]))
    integ, prev = np.zeros(2), np.zeros(2)
    rng = np.random.default_rng(0)
    for _ in range(20):
        s = rng.normal(size=4).astype(np.float32)
        e = np.array([1.0, -1.0]) - s[:2]
        integ += e
        want = np.clip(0.5 * e + 0.1 * integ + 0.05 * (e - prev), -1, 1)
        prev = e
        assert np.array_equal(ag.act(s), want)
    _, pp = ag.device_policy()
    assert (pp.baseline.kp, pp.baseline.ki, pp.baseline.kd) == (0.5, 0.1, 0.05) and list(pp.baseline.setpoint)[:2] == [1.0, -1.0]
