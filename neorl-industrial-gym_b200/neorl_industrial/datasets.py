"""get_dataset-style generation on the device (reference chemical_reactor.py:324-420, power_grid.py:194-249,
robot_assembly.py:246-308): every episode is an independent env; a length-probe pass, an exclusive scan
and a write pass emit episode-contiguous D4RL arrays straight into HBM, exported with pinned async copies."""
from __future__ import annotations

import os
from typing import Dict, Optional

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
    """The smallest episode count whose dataset holds >= n_transitions rows (so the overshoot is less than one episode).
    Episodes are independent envs keyed by their index, so the row count is monotone in the episode count: grow
    geometrically until the target is reached, then bisect. Each probe is one length-only pass (nig_dataset_size); the
    probe of the final count is cached by the library and not repeated by the write pass."""
    if n_transitions <= 0:
        raise ValueError("n_transitions must be positive")
    lo, hi = 0, max(1, -(-n_transitions // n_steps))       # lo: too few (0 episodes = 0 rows); hi: candidate
    for _ in range(32):
        m = native.dataset_size(hi, n_steps, policy, params)
        if m >= n_transitions:
            break
        lo, hi = hi, max(hi + 1, int(hi * (n_transitions / max(m, 1)) * 1.02) + 1)
    else:
        raise RuntimeError("could not reach the requested number of transitions (episodes end immediately?)")
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if native.dataset_size(mid, n_steps, policy, params) >= n_transitions:
            hi = mid
        else:
            lo = mid
    native.dataset_size(hi, n_steps, policy, params)       # leave the final count as the cached probe
    return hi


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


# ------------------------------------------------------------------------------------------------------------------
# Dataset containers and file formats. The reference keeps datasets in memory only; its docs promise HDF5 export
# (docs/ARCHITECTURE.md:65,77; h5py in pyproject.toml:36) and its test fixture (tests/fixtures/industrial_data.py:11-21)
# uses the 8-field layout below. Host-side plumbing: nothing here touches the step path.
FIXTURE_FIELDS = ("observations", "actions", "rewards", "next_observations", "terminals", "timeouts", "safety_violations")


def to_fixture_layout(ds: Dict[str, np.ndarray], metadata: Optional[dict] = None) -> Dict[str, object]:
    """A get_dataset(..., extensions=True) dict in the IndustrialDataset layout of tests/fixtures/industrial_data.py:
    observations, actions, rewards, next_observations, terminals, timeouts, safety_violations (bool), metadata."""
    if "next_observations" not in ds or "safety" not in ds:
        raise KeyError("the fixture layout needs next_observations and safety: call get_dataset(..., extensions=True)")
    m = int(ds["rewards"].shape[0])
    out = {k: ds[k] for k in ("observations", "actions", "rewards", "next_observations", "terminals")}
    out["timeouts"] = ds["timeouts"] if "timeouts" in ds else np.zeros(m, bool)
    out["safety_violations"] = np.asarray(ds["safety"]) != 0
    out["metadata"] = dict(metadata or {})
    out["metadata"].setdefault("n_transitions", m)
    return out


def save_dataset(path: str, ds: Dict[str, object], metadata: Optional[dict] = None) -> str:
    """Write a dataset dict to ``.npz`` (numpy, compressed), ``.h5`` / ``.hdf5`` (needs h5py) or ``.pt`` (torch.save).
    ``metadata`` (or ds['metadata']) is stored as a JSON string / HDF5 attributes."""
    import json
    meta = dict(ds.get("metadata", {}) or {})
    meta.update(metadata or {})
    arrays = {k: np.asarray(v) for k, v in ds.items() if k != "metadata"}
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        np.savez_compressed(path, __metadata__=np.array(json.dumps(meta)), **arrays)
    elif ext in (".h5", ".hdf5"):
        try:
            import h5py
        except ImportError as e:
            raise ImportError("HDF5 export needs h5py, which is not installed; use .npz or .pt") from e
        with h5py.File(path, "w") as f:
            for k, v in arrays.items():
                f.create_dataset(k, data=v, compression="gzip")
            for k, v in meta.items():
                f.attrs[k] = json.dumps(v)
    elif ext == ".pt":
        import torch
        torch.save({"metadata": meta, **{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in arrays.items()}}, path)
    else:
        raise ValueError(f"unknown dataset file extension {ext!r}; expected .npz, .h5/.hdf5 or .pt")
    return path


def load_dataset(path: str) -> Dict[str, object]:
    """Inverse of save_dataset: dict of numpy arrays + 'metadata'."""
    import json
    ext = os.path.splitext(path)[1].lower()
    if ext == ".npz":
        with np.load(path, allow_pickle=False) as z:
            out = {k: z[k] for k in z.files if k != "__metadata__"}
            out["metadata"] = json.loads(str(z["__metadata__"])) if "__metadata__" in z.files else {}
        return out
    if ext in (".h5", ".hdf5"):
        try:
            import h5py
        except ImportError as e:
            raise ImportError("reading HDF5 needs h5py, which is not installed") from e
        with h5py.File(path, "r") as f:
            out = {k: f[k][()] for k in f.keys()}
            out["metadata"] = {k: json.loads(v) for k, v in f.attrs.items()}
        return out
    if ext == ".pt":
        import torch
        d = torch.load(path, weights_only=False)
        return {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in d.items()}
    raise ValueError(f"unknown dataset file extension {ext!r}")


def as_torch(ds: Dict[str, object], device=None, pin_memory: bool = False):
    """Torch-tensor view of a dataset dict for offline-RL data loaders (zero-copy from numpy on the host; ``device``
    moves it). bool arrays stay bool."""
    import torch
    out = {}
    for k, v in ds.items():
        if k == "metadata":
            out[k] = v
            continue
        t = torch.from_numpy(np.ascontiguousarray(v))
        if pin_memory:
            t = t.pin_memory()
        out[k] = t.to(device, non_blocking=pin_memory) if device is not None else t
    return out
