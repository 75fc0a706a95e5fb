"""ctypes binding of the CPU oracle (oracle/nig_oracle.c). TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Importable only from tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs.
The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libnig_oracle.so")

REACTOR, GRID, ROBOT = 0, 1, 2
KIND_BY_ID = {"ChemicalReactor-v0": REACTOR, "PowerGrid-v0": GRID, "RobotAssembly-v0": ROBOT}
CON_BUILTIN, CON_BOUND, CON_HOSTMASK = 0, 1, 2
F_TERMINATED, F_TRUNCATED, F_CRITICAL, F_RESET, F_INACTIVE = 1, 2, 4, 8, 128
MAX_CONS = 8
STATE_DIM = {REACTOR: 12, GRID: 32, ROBOT: 24}
ACTION_DIM = {REACTOR: 3, GRID: 8, ROBOT: 7}
NOISE_DIM = {REACTOR: 2, GRID: 23, ROBOT: 0}
MAX_EPISODE_STEPS = {REACTOR: 500, GRID: 1000, ROBOT: 1000}
# (penalty, critical) of the built-in constraints, in registration order
BUILTIN_CONS = {
    REACTOR: [(-100.0, 1), (-50.0, 1), (-25.0, 0)],   # chemical_reactor.py:38-60
    GRID: [(-50.0, 1), (-30.0, 1), (-20.0, 0)],       # power_grid.py:53-72
    ROBOT: [(-100.0, 1), (-200.0, 1), (-50.0, 0)],    # robot_assembly.py:56-75
}


class Con(C.Structure):
    _fields_ = [("kind", C.c_int32), ("id", C.c_int32), ("si", C.c_int32), ("ai", C.c_int32),
                ("coef", C.c_float), ("lo", C.c_float), ("hi", C.c_float), ("penalty", C.c_float),
                ("critical", C.c_int32)]


class Cfg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("max_episode_steps", C.c_int32), ("n_cons", C.c_int32),
                ("exp_mode", C.c_int32), ("auto_reset", C.c_int32), ("pad", C.c_int32),
                ("seed", C.c_uint64), ("cons", Con * MAX_CONS)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nig_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnig_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        assert _lib.orc_cfg_size() == C.sizeof(Cfg)
        _lib.orc_spec_expf.restype = C.c_float
        _lib.orc_spec_expf.argtypes = [C.c_float]
        _lib.orc_spec_normal.restype = C.c_float
        _lib.orc_spec_normal.argtypes = [C.c_uint32]
    return _lib


def _p(a, ty=None):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def make_cfg(kind, *, exp_mode=0, auto_reset=True, seed=0, max_episode_steps=None, extra_cons=(), builtin=True):
    cfg = Cfg()
    cfg.kind = kind
    cfg.max_episode_steps = MAX_EPISODE_STEPS[kind] if max_episode_steps is None else max_episode_steps
    cfg.exp_mode = exp_mode
    cfg.auto_reset = int(auto_reset)
    cfg.seed = seed
    k = 0
    if builtin:
        for i, (pen, crit) in enumerate(BUILTIN_CONS[kind]):
            cfg.cons[k] = Con(CON_BUILTIN, i, 0, -1, 0.0, 0.0, 0.0, pen, crit)
            k += 1
    for c in extra_cons:
        cfg.cons[k] = c
        k += 1
    cfg.n_cons = k
    return cfg


def bound_con(si, lo, hi, penalty, critical=False, ai=-1, coef=0.0):
    return Con(CON_BOUND, 0, si, ai, coef, lo, hi, penalty, int(critical))


def hostmask_con(bit, penalty, critical=False):
    return Con(CON_HOSTMASK, bit, 0, -1, 0.0, 0.0, 0.0, penalty, int(critical))


class PolicyParams(C.Structure):
    _fields_ = [("p_ctrl", C.c_float), ("uniform_scale", C.c_float), ("store_clip", C.c_float), ("mode", C.c_int32),
                ("gain", (C.c_float * 2) * 8), ("sigma", C.c_float * 8)]


POLICY_ACTIONS, POLICY_UNIFORM, POLICY_ZERO, POLICY_PCTRL = 0, 1, 2, 3


def copy_policy_params(src) -> "PolicyParams":
    """Field-wise copy from any ctypes struct with the same field names (e.g. the product's PolicyParams)."""
    pp = PolicyParams()
    pp.p_ctrl, pp.uniform_scale, pp.store_clip, pp.mode = src.p_ctrl, src.uniform_scale, src.store_clip, src.mode
    for k in range(8):
        pp.gain[k][0], pp.gain[k][1], pp.sigma[k] = src.gain[k][0], src.gain[k][1], src.sigma[k]
    return pp


class OracleEnv:
    """Batched AoS env state stepped by the C oracle (mirrors IndustrialEnv.step, base.py:157-213)."""

    def __init__(self, kind, n, *, env_id0=0, threads=1, **cfg_kw):
        self.kind, self.n, self.env_id0, self.threads = kind, int(n), int(env_id0), threads
        self.cfg = make_cfg(kind, **cfg_kw)
        self.S, self.A, self.NZ = STATE_DIM[kind], ACTION_DIM[kind], NOISE_DIM[kind]
        self.state = np.zeros((self.n, self.S), np.float32)
        self.ep_step = np.zeros(self.n, np.int32)
        self.ep_viol = np.zeros(self.n, np.int32)
        self.done_latch = np.zeros(self.n, np.uint8)
        self.stats = np.zeros(16, np.int64)
        self.tick = 0
        self.epoch = 0

    def reset(self, mask=None, init_states=None):
        if init_states is not None:
            init_states = np.ascontiguousarray(init_states, np.float32)
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
        self.epoch += 1   # explicit resets draw with a fresh epoch; auto-resets reuse the current one
        lib().orc_reset_batch(C.byref(self.cfg), C.c_int64(self.n), C.c_int64(self.env_id0),
                              C.c_uint32(self.tick), C.c_uint32(self.epoch),
                              _p(self.state), _p(self.ep_step), _p(self.ep_viol), _p(self.done_latch),
                              _p(mask), _p(init_states))
        return self.state.copy()

    def step(self, actions, noise=None, reset_states=None, hostmask=None, want_next_obs=True):
        actions = np.ascontiguousarray(actions, np.float32).reshape(self.n, self.A)
        if noise is not None:
            noise = np.ascontiguousarray(noise, np.float32).reshape(self.n, self.NZ)
        if reset_states is not None:
            reset_states = np.ascontiguousarray(reset_states, np.float32).reshape(self.n, self.S)
        if hostmask is not None:
            hostmask = np.ascontiguousarray(hostmask, np.uint8)
        next_obs = np.empty((self.n, self.S), np.float32) if want_next_obs else None
        reward = np.empty(self.n, np.float32)
        flags = np.empty(self.n, np.uint8)
        viol = np.empty(self.n, np.uint8)
        lib().orc_step_batch(C.byref(self.cfg), C.c_int64(self.n), C.c_int64(self.env_id0),
                             C.c_uint32(self.tick), C.c_uint32(self.epoch),
                             _p(self.state), _p(self.ep_step), _p(self.ep_viol), _p(self.done_latch),
                             _p(actions), _p(noise), _p(reset_states), _p(hostmask),
                             _p(next_obs), _p(reward), _p(flags), _p(viol), _p(self.stats),
                             C.c_int(self.threads))
        self.tick += 1
        return next_obs, reward, flags, viol


def rollout(env: "OracleEnv", n_steps, policy, pp=None, want_reward_sum=False):
    """n_steps free-running steps inside one C call (threads own env blocks for the whole horizon)."""
    pp = pp if pp is not None else PolicyParams()
    rs = np.zeros(env.n, np.float32) if want_reward_sum else None
    lib().orc_rollout_batch(C.byref(env.cfg), C.c_int(policy), C.byref(pp), C.c_int64(env.n), C.c_int64(env.env_id0),
                            C.c_uint32(env.tick), C.c_uint32(env.epoch), C.c_int32(n_steps),
                            _p(env.state), _p(env.ep_step), _p(env.ep_viol), _p(env.done_latch), _p(rs), _p(env.stats),
                            C.c_int(env.threads))
    env.tick += n_steps
    return rs


def policy_actions(env: "OracleEnv", policy, pp=None):
    """Actions the in-kernel policy produces for every env at env.tick (before stepping)."""
    a = np.empty((env.n, env.A), np.float32)
    pp = pp if pp is not None else PolicyParams()
    lib().orc_policy_batch(C.byref(env.cfg), C.c_int(policy), C.byref(pp), C.c_int64(env.n), C.c_int64(env.env_id0),
                           C.c_uint32(env.tick), _p(env.state), _p(a), C.c_int(env.threads))
    return a


def policy_forced(kind, pp, state, coin=None, z=None, u=None):
    """get_dataset's policy arithmetic on supplied random inputs (coin [n], z [n, 8] normals, u [n, 8] in [-1, 1])."""
    state = np.ascontiguousarray(state, np.float32)
    n = state.shape[0]
    f32 = lambda x: None if x is None else np.ascontiguousarray(x, np.float32)
    coin, z, u = f32(coin), f32(z), f32(u)
    a = np.empty((n, ACTION_DIM[kind]), np.float32)
    cfg = make_cfg(kind)
    lib().orc_policy_forced(C.byref(cfg), C.byref(pp), C.c_int64(n), _p(state), _p(coin), _p(z), _p(u), _p(a))
    return a


def dynamics(kind, s, a, nz=None, exp_mode=0):
    s = np.ascontiguousarray(s, np.float32)
    a = np.ascontiguousarray(a, np.float32)
    n = s.shape[0]
    if nz is None:
        nz = np.zeros((n, max(NOISE_DIM[kind], 1)), np.float32)
    nz = np.ascontiguousarray(nz, np.float32)
    o = np.empty_like(s)
    lib().orc_dynamics(C.c_int(kind), C.c_int(exp_mode), C.c_int64(n), _p(s), _p(a), _p(nz), _p(o))
    return o


def reward(kind, ns, a):
    ns = np.ascontiguousarray(ns, np.float32)
    a = np.ascontiguousarray(a, np.float32)
    r = np.empty(ns.shape[0], np.float64)
    lib().orc_reward(C.c_int(kind), C.c_int64(ns.shape[0]), _p(ns), _p(a), _p(r))
    return r


def is_done(kind, s):
    s = np.ascontiguousarray(s, np.float32)
    d = np.empty(s.shape[0], np.uint8)
    lib().orc_is_done(C.c_int(kind), C.c_int64(s.shape[0]), _p(s), _p(d))
    return d.astype(bool)


def spec_normals4(seed, env, tick, stream, j):
    z = np.empty(4, np.float32)
    lib().orc_spec_normals4(C.c_uint64(seed), C.c_uint32(env), C.c_uint32(tick), C.c_uint32(stream), C.c_uint32(j), _p(z))
    return z


def spec_normal(w) -> float:
    """The spec's standard normal of one 32-bit word (inverse CDF, dyadic-segment table)."""
    return float(lib().orc_spec_normal(C.c_uint32(int(w) & 0xFFFFFFFF)))


def selftest_normal(first, stride, count):
    """The two wrapping checksums nig_selftest_normal computes on the device (include/nig_b200.h)."""
    out = np.zeros(2, np.uint64)
    lib().orc_selftest_normal(C.c_uint32(first), C.c_uint32(stride), C.c_int64(count), _p(out))
    return int(out[0]), int(out[1])


def philox(c, k, rounds=10):
    out = np.empty(4, np.uint32)
    lib().orc_philox(C.c_int(rounds), *(C.c_uint32(int(x)) for x in c), *(C.c_uint32(int(x)) for x in k), _p(out))
    return out


def words_batch(seed, env0, n_env, tick0, n_tick, stream=0, j=0):
    """uint32 [n_env, n_tick, 4]: the spec's Philox words of a rectangle of (env, tick) counters."""
    out = np.empty((n_env, n_tick, 4), np.uint32)
    lib().orc_words_batch(C.c_uint64(seed), C.c_uint32(env0), C.c_int32(n_env), C.c_uint32(tick0), C.c_int32(n_tick),
                          C.c_uint32(stream), C.c_uint32(j), _p(out))
    return out


def max_threads() -> int:
    return int(lib().orc_max_threads())
