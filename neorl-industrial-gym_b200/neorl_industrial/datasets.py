"""get_dataset-style generation on the device (reference chemical_reactor.py:324-420, power_grid.py:194-249,
robot_assembly.py:246-308): every episode is an independent env; a length-probe pass, an exclusive scan
and a write pass emit episode-contiguous D4RL arrays straight into HBM, exported with pinned async copies."""
from __future__ import annotations

from typing import Dict

import numpy as np

from . import _native as N


def generate_dataset_device(native, n_episodes: int, n_steps: int, policy: int, params, *, extensions: bool = False,
                            timeouts: bool = True, terminals_include_truncation: bool = True):
    """Returns (dict of torch CUDA tensors, n_transitions). Layout: observations [M,S] f32, actions [M,A] f32,
    rewards [M] f32, terminals [M] u8, timeouts [M] u8 (+ next_observations [M,S], safety [M] u8)."""
    import torch
    dev = native.torch_device()
    m = native.dataset_size(n_episodes, n_steps, policy, params)
    cap = max(m, 1)
    out = {
        "observations": torch.empty((cap, native.S), dtype=torch.float32, device=dev),
        "actions": torch.empty((cap, native.A), dtype=torch.float32, device=dev),
        "rewards": torch.empty((cap,), dtype=torch.float32, device=dev),
        "terminals": torch.empty((cap,), dtype=torch.uint8, device=dev),
    }
    if timeouts:
        out["timeouts"] = torch.empty((cap,), dtype=torch.uint8, device=dev)
    if extensions:
        out["next_observations"] = torch.empty((cap, native.S), dtype=torch.float32, device=dev)
        out["safety"] = torch.empty((cap,), dtype=torch.uint8, device=dev)
    written = native.dataset_device(n_episodes, n_steps, policy, params, out, cap, terminals_include_truncation)
    assert written == m, (written, m)
    return {k: v[:m] for k, v in out.items()}, m


def episodes_for_transitions(native, n_transitions: int, n_steps: int, policy: int, params) -> int:
    """Smallest episode count found whose dataset holds >= n_transitions rows. Episodes are independent envs keyed by
    their index, so the row count is monotone in the episode count; the length-probe pass (nig_dataset_size) of the
    final count is cached by the library and not repeated by the write pass."""
    if n_transitions <= 0:
        raise ValueError("n_transitions must be positive")
    n_ep = max(1, -(-n_transitions // n_steps))
    for _ in range(32):
        m = native.dataset_size(n_ep, n_steps, policy, params)
        if m >= n_transitions:
            return n_ep
        n_ep = max(n_ep + 1, int(n_ep * (n_transitions / max(m, 1)) * 1.02) + 1)
    raise RuntimeError("could not reach the requested number of transitions (episodes end immediately?)")


def generate_dataset(env, n_episodes: int, n_steps: int, policy: int, params, *, terminals_include_truncation: bool,
                     timeouts_key: bool, extensions: bool = False) -> Dict[str, np.ndarray]:
    """Host dict with the reference's keys and dtypes (bool terminals / timeouts)."""
    import torch
    dev_out, m = generate_dataset_device(env.native, n_episodes, n_steps, policy, params, extensions=extensions,
                                         timeouts=timeouts_key, terminals_include_truncation=terminals_include_truncation)
    host = {}
    for k, v in dev_out.items():
        pinned = torch.empty(v.shape, dtype=v.dtype, pin_memory=True)
        pinned.copy_(v, non_blocking=True)
        host[k] = pinned
    torch.cuda.synchronize(env.native.torch_device())
    res = {k: v.numpy().copy() for k, v in host.items()}
    for k in ("terminals", "timeouts"):
        if k in res:
            res[k] = res[k].astype(bool)
    return res
