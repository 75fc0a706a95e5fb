"""What the reference's own test-suite asserts about the real envs (tests/test_comprehensive_system.py:24-110 and
:585-625 upstream), asserted here against the drop-in package: a user's existing checks keep passing."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ENV_DIMS = [("ChemicalReactor-v0", 12, 3), ("PowerGrid-v0", 32, 8), ("RobotAssembly-v0", 24, 7)]


@pytest.fixture(scope="module")
def ni():
    import neorl_industrial as ni
    return ni


def test_every_registered_env_can_be_made(ni):
    for env_id, s, a in ENV_DIMS:
        env = ni.make(env_id)
        assert env is not None and env.state_dim == s and env.action_dim == a
        env.close()


def test_episode_lifecycle_through_the_gym_api(ni):
    env = ni.make("ChemicalReactor-v0")
    obs, info = env.reset()
    assert obs.shape == (env.state_dim,) and isinstance(info, dict)
    steps, done, total = 0, False, 0.0
    while not done and steps < 100:
        nxt, reward, terminated, truncated, info = env.step(env.action_space.sample())
        assert nxt.shape == (env.state_dim,)
        assert isinstance(reward, (int, float, np.floating))       # upstream returns np.float32 for this env too
        assert isinstance(terminated, bool) and isinstance(truncated, bool) and isinstance(info, dict)
        total += float(reward); steps += 1
        done = terminated or truncated
    assert steps > 0 and abs(total) < 1e6
    env.close()


def test_constraint_check_fns_are_callable_on_observations(ni):
    env = ni.make("ChemicalReactor-v0")
    obs, _ = env.reset()
    action = env.action_space.sample()
    assert len(env.safety_constraints) > 0
    for c in env.safety_constraints:
        assert isinstance(c.check_fn(obs, action), (bool, np.bool_))
    env.close()


@pytest.mark.parametrize("cls_name", ["ChemicalReactorEnv", "PowerGridEnv", "RobotAssemblyEnv"])
def test_dataset_generation_for_every_env_and_quality(ni, cls_name):
    env = getattr(ni, cls_name)()
    for quality in ("expert", "medium", "mixed", "random"):
        ds = env.get_dataset(quality=quality)
        for key in ("observations", "actions", "rewards", "terminals"):
            assert key in ds and len(ds[key]) > 0
        assert ds["observations"].shape[1] == env.state_dim and ds["actions"].shape[1] == env.action_dim
        assert len(ds["observations"]) == len(ds["actions"]) == len(ds["rewards"])
    env.close()


def test_one_interaction_per_env_and_safety_metrics_in_info(ni):
    for env_id, s, a in ENV_DIMS:
        env = ni.make(env_id)
        obs, _ = env.reset()
        action = env.action_space.sample()
        nxt, reward, terminated, truncated, info = env.step(action)
        assert obs.shape == (s,) and nxt.shape == (s,) and len(action) == a
        sm = info["safety_metrics"]
        assert hasattr(sm, "violation_count") and hasattr(sm, "safety_score")
        env.close()
