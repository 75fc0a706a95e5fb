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


@pytest.mark.parametrize("name,policy", [("reactor", "random"), ("grid", "random"), ("reactor", "pid")])
def test_rollout_host_env_slices_equal_one_launch_sequence(mods, name, policy, monkeypatch):
    """nig_rollout_host splits a large population into env slices on separate streams (copies of one slice overlap the
    stepping of the others). Every output must equal the unsliced call's bit for bit -- trajectories are keyed by global
    env id and tick -- for a ragged last slice, initial states from the host, two consecutive calls, and the
    stateful in-kernel PID controller."""
    ni, N, O = mods
    n, T, K = 3 * 8192 + 1000, 130, 64
    cls = {"reactor": ni.ChemicalReactorEnv, "grid": ni.PowerGridEnv}[name]
    out = {}
    init = None
    for slices in (1, 3):
        monkeypatch.setenv("NIG_HOST_SLICES", str(slices))
        env = cls(num_envs=n, seed=33, env_id_offset=7)
        if policy == "pid":
            from neorl_industrial.benchmarks.baseline_agents import PIDControllerAgent
            policy = PIDControllerAgent(env.state_dim, env.action_dim, kp=0.01, ki=0.001, kd=0.001)
        drawn = env.native.reset_host()                  # (every run: explicit resets advance the epoch that keys the reset draws)
        if init is None:
            init = drawn.copy()
            init[:, 0] += np.float32(0.25)
        a = env.rollout(T, policy, steps_per_launch=K, init_states=init)
        b = env.rollout(40, policy, steps_per_launch=K)                   # continues without a reset
        out[slices] = (a, b, env.native.launch_count)
        env.close()
    for x, y in zip(out[1][:2], out[3][:2]):
        for key in ("obs", "reward_sum", "violations", "episodes"):
            assert_bits_equal(x[key], y[key], f"{key} sliced vs unsliced")
        for key in ("steps", "episodes", "terminated", "truncated", "critical_shutdowns", "violations", "successes",
                    "episode_length_sum", "violations_per_constraint"):
            assert x["stats"][key] == y["stats"][key], key
        np.testing.assert_allclose(x["stats"]["return_sum"], y["stats"]["return_sum"], rtol=1e-12)
    assert out[3][2] > out[1][2]                                           # the sliced calls really launched per slice


@pytest.mark.parametrize("slices", [1, 3])
def test_rollout_steps_device_equals_launch_sequence(mods, slices, monkeypatch):
    """nig_rollout_steps (T steps as fused K-step launches, env slices on internal streams) == the same launches issued one
    by one with nig_rollout on one stream: state, episode counters, per-env outputs, stats."""
    import torch
    ni, N, O = mods
    monkeypatch.setenv("NIG_HOST_SLICES", str(slices))
    n, T, K = 3 * 8192 + 1000, 150, 64
    a = ni.NativeEnv(N.ENV_CHEMICAL_REACTOR, n, device=0, seed=9, env_id_offset=3)
    b = ni.NativeEnv(N.ENV_CHEMICAL_REACTOR, n, device=0, seed=9, env_id_offset=3)
    a.reset_host(); b.reset_host()
    rs, vc, dc = a.empty(), a.empty(dtype=torch.int32), a.empty(dtype=torch.int32)
    a.rollout_steps_device(T, K, N.POLICY_UNIFORM, reward_sum=rs, viol_count=vc, done_count=dc)
    a.rollout_steps_device(20, K, N.POLICY_UNIFORM, reward_sum=rs, viol_count=vc, done_count=dc, accumulate=True)
    ref_rs = np.zeros(n, np.float32); ref_vc = np.zeros(n, np.int64); ref_dc = np.zeros(n, np.int64)
    r1, v1, d1 = b.empty(), b.empty(dtype=torch.int32), b.empty(dtype=torch.int32)
    for c, k in enumerate((64, 64, 22, 20)):
        b.rollout_device(k, N.POLICY_UNIFORM, reward_sum=r1, viol_count=v1, done_count=d1)
        torch.cuda.synchronize()
        part = r1[:n].cpu().numpy()
        ref_rs = part.copy() if c == 0 else (ref_rs + part).astype(np.float32)
        ref_vc += v1[:n].cpu().numpy(); ref_dc += d1[:n].cpu().numpy()
    torch.cuda.synchronize()
    for x, y, what in zip(a.get_state_host(), b.get_state_host(), ("state", "ep_step", "ep_viol", "done")):
        assert_bits_equal(x, y, what)
    assert_bits_equal(rs[:n].cpu().numpy(), ref_rs, "reward_sum")
    assert np.array_equal(vc[:n].cpu().numpy(), ref_vc) and np.array_equal(dc[:n].cpu().numpy(), ref_dc)
    sa, sb = a.stats_dict(), b.stats_dict()
    for key in ("steps", "episodes", "terminated", "truncated", "violations", "successes", "episode_length_sum", "violations_per_constraint"):
        assert sa[key] == sb[key], key
    assert sa["steps"] == n * (T + 20)
    with pytest.raises(NotImplementedError):
        r = N.Rollout(); r.n_steps, r.policy = 8, N.POLICY_ACTIONS
        import ctypes as C
        N.check(N.lib().nig_rollout_steps(a._h, C.byref(r), 16, None))
    a.close(); b.close()


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


@pytest.mark.parametrize("name", ["reactor", "grid", "robot"])
def test_rollout_host_direct_mode_equals_copy_path(name, monkeypatch):
    """nig_rollout_host has two data paths for the caller's host arrays: direct (the slices' ingest / export kernels read and
    write the page-locked arrays over PCIe themselves) and staged (cudaMemcpyAsync through device buffers). Which one runs
    depends on NIG_HOST_DIRECT and on every supplied array being page-locked and 16-byte aligned; the results may not.
    Four calls through the raw C ABI -- pinned + aligned (direct), NIG_HOST_DIRECT=0, pageable numpy arrays, pinned arrays
    shifted by 4 bytes -- from the same initial states: bit-identical outputs. n is not a multiple of 128 (ragged last slice)."""
    import ctypes as C
    import neorl_industrial as ni
    from neorl_industrial import _native as N
    kind = KINDS[name]
    n, T, K = 33_000 + 77, 150, 64
    rng = np.random.default_rng(4)

    def call(direct, how):
        monkeypatch.setenv("NIG_HOST_DIRECT", "3" if direct else "0")          # bit 0: ingest, bit 1: export
        monkeypatch.setenv("NIG_HOST_DIRECT_MAX_MB", "1e9")                      # no size thresholds: both sides direct
        env = ni.NativeEnv(kind, n, device=0, seed=91)
        S = env.S
        init0 = env.reset_host().copy()
        keep = []

        guards = []

        def host(shape, dtype):
            cnt = int(np.prod(shape))
            if how == "pageable":
                full = np.zeros(cnt + 4, dtype)
            else:
                p = N.PinnedArray((cnt + 4,), dtype)
                keep.append(p)
                full = p.array
            lo = 1 if how == "shifted" else 0
            full.view(np.uint32)[:] = 0xA5A5A5A5                   # canaries around the exact-size window the library may touch
            guards.append((full, lo, cnt))
            return full[lo:lo + cnt].reshape(shape)
        init, obs = host((n, S), np.float32), host((n, S), np.float32)
        rew, vi, dn = host((n,), np.float32), host((n,), np.int32), host((n,), np.int32)
        init[:] = init0
        counters, sums = np.zeros(24, np.int64), np.zeros(8, np.float64)
        r = N.RolloutHost()
        r.n_steps, r.steps_per_launch, r.policy, r.reset_first = T, K, N.POLICY_UNIFORM, 0
        r.init_states, r.final_obs = N.ptr_of(init), N.ptr_of(obs)
        r.reward_sum, r.viol_count, r.done_count = N.ptr_of(rew), N.ptr_of(vi), N.ptr_of(dn)
        r.counters24, r.sums8 = N.ptr_of(counters), N.ptr_of(sums)
        for _ in range(2):                                   # second call = graph replay where a graph is used
            N.check(N.lib().nig_rollout_host(env._h, C.byref(r)))
        res = (obs.copy(), rew.copy(), vi.copy(), dn.copy(), counters.copy(), sums.copy())
        for full, lo, cnt in guards:
            w = full.view(np.uint32)
            assert (w[:lo] == 0xA5A5A5A5).all() and (w[lo + cnt:] == 0xA5A5A5A5).all(), "bytes outside a host array were overwritten"
        env.close()
        return res

    ref = call(True, "pinned")
    assert ref[4][0] == 2 * n * T
    for direct, how in ((False, "pinned"), (True, "pageable"), (True, "shifted")):
        got = call(direct, how)
        assert_bits_equal(got[0], ref[0], f"{how} direct={direct}: observations")
        assert_bits_equal(got[1], ref[1], f"{how} direct={direct}: reward sums")
        assert np.array_equal(got[2], ref[2]) and np.array_equal(got[3], ref[3])
        assert got[4].tolist() == ref[4].tolist()
        np.testing.assert_allclose(got[5], ref[5], rtol=1e-12)      # fp64 sums: the order of the CTAs' atomic adds is not fixed
