"""Latency of the single-env gym API (BASELINE config 1 shape): env.step() and the bare C call, zero-copy on / off."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neorl-industrial-gym_b200")]
import numpy as np
import neorl_industrial as ni
for zc in ("1", "0"):
    os.environ["NIG_ZERO_COPY"] = zc
    env = ni.make("ChemicalReactor-v0"); env.reset()
    a = env.action_space.sample()
    for _ in range(200): env.step(a)
    t0 = time.time()
    for _ in range(3000):
        o, r, te, tr, i = env.step(a)
        if te or tr: env.reset()
    full = (time.time() - t0) / 3000 * 1e6
    nat, a1 = env.native, a.reshape(1, 3)
    t0 = time.time()
    for _ in range(3000): nat.step_host(a1)
    bare = (time.time() - t0) / 3000 * 1e6
    print(f"zero_copy={zc}: env.step {full:.1f} us ({1e6/full:.0f} steps/s), native.step_host {bare:.1f} us", flush=True)
    env.close()
